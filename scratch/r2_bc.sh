#!/bin/bash
# long-transform kernels (longconv.cu k_bc_*): parity (new tests, every convolution test, the isolating variants),
# then cfg4 A/B: generic three-kernel path against the c2 kernels, each alone and together
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_convolution or convolution_voltages or cfg4 or one_kernel" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "long_" 2>&1 | tail -4
L=$PWD/dspsr_b200/libb200dsp_dev.so
run() {  # tag env...
  tag=$1; shift
  env B200_LIB=$L "$@" python bench.py --workload cfg4 --steps 6 --warmup 3 --no-cpu > gpurun_out/r2s/bench_bc_$tag.json 2> gpurun_out/r2s/bench_bc_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s/bench_bc_$tag.json"))
    print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_block"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2s/bench_bc_$tag.err").read()[-800:])
PY
}
run generic B200_BIG_CONV=0
run all B200_BIG_CONV=1
python bench.py --workload cfg4 --steps 6 --warmup 3 --no-cpu > gpurun_out/r2s/bench_cfg4_product.json 2> gpurun_out/r2s/bench_cfg4_product.err; tail -c 600 gpurun_out/r2s/bench_cfg4_product.json
