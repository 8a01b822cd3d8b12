// libvsl.so is one translation unit: the fused path launches kernels defined beside the stand-alone ops.
#include "vsl_ops.cu"
#include "vsl_loss.cu"
