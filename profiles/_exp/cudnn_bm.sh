#!/bin/bash
for gb in 256 32; do for bm in 0 1; do
  echo -n "gb=$gb cudnn.benchmark=$bm: "
  VSL_CUDNN_BENCHMARK=$bm python profiles/train_samples.py --global-batch $gb 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("ms", round(d["ms_per_step"],2), "samples/s", round(d["value"],1), "graph", d["cuda_graph"])'
done; done
