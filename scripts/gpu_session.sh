#!/bin/bash
# Development tool: one GPU-box session -- GPU tests, both bench arms, kernel-variant timings, PCIe probe and the ncu captures.
# Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
if [ ! -f tests/golden/tcnn_loss1000_hash_ob_d6.npz ]; then
  timeout 900 python tests/golden/make_tcnn_loss_curve.py > gpurun_out/loss_gen.log 2>&1 && cp gpurun_out/tcnn_loss1000_*.npz tests/golden/
fi
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 120 python scripts/pcie_probe.py > gpurun_out/pcie.json 2>&1; cat gpurun_out/pcie.json
timeout 600 python bench.py --impl reference --steps 100 --warmup 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"
cut -c1-1500 gpurun_out/bench_ours.json
: > gpurun_out/tune.log
timeout 300 python scripts/tune.py 0,1 >> gpurun_out/tune.log 2>&1
for v in nrc_hpm_renderer_b200/variants/*.so; do
  [ -f "$v" ] || continue
  echo "== $v" >> gpurun_out/tune.log
  NRCHPM_LIB=$PWD/$v timeout 300 python scripts/tune.py 0,1 >> gpurun_out/tune.log 2>&1
done
grep -v Traceback gpurun_out/tune.log | cut -c1-220
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 --no-frame > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nrc_forward_kernel -s 6 -c 1 -o gpurun_out/prof_fwd_infer python bench.py --steps 4 --warmup 3 --no-frame > gpurun_out/ncu_full.log 2>&1
echo done
