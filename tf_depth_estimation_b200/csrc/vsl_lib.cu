// libvsl.so is one translation unit: the fused path launches kernels defined beside the stand-alone ops.
#include "vsl_ops.cu"
#include "vsl_loss.cu"
#include "vsl_ext.cu"
#include "vsl_optim.cu"
