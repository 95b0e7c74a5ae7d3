#!/bin/bash
# Round-2 ncu captures on ONE GPU (B200_PROFILING.md recipe).  Outputs in gpurun_out/, distilled into profiles/ by
# scripts/ncu_summary.py.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none --launch-count 1"
SW="python -m oktopk_b200.bench.sweep --schemes oktopk --iters 6 --warmup 3"
# (1) every launch of three eager VGG-16 steps with its device time (shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1200 --csv --log-file gpurun_out/launches_vgg.csv \
    python bench.py --steps 3 --warmup 4 --dense-warmup 4 --no-e2e --no-extra --no-graph > gpurun_out/launches_vgg.log 2>&1
# (2) the fused Ok-Topk kernel, VGG-16-sized bucket, density 0.001, steady-state call
$NCU --kernel-name regex:oktopk_fused --launch-skip 6 -f -o gpurun_out/okt_vgg_d001 $SW --sizes 14728266 --densities 0.001 > gpurun_out/ncu_okt_vgg.log 2>&1
# (3) 128 Mi bucket at density 0.001 / 0.01
$NCU --kernel-name regex:oktopk_fused --launch-skip 6 -f -o gpurun_out/okt_128m_d001 $SW --sizes 128M --densities 0.001 > gpurun_out/ncu_okt_128m_d001.log 2>&1
$NCU --kernel-name regex:oktopk_fused --launch-skip 6 -f -o gpurun_out/okt_128m_d01 $SW --sizes 128M --densities 0.01 > gpurun_out/ncu_okt_128m_d01.log 2>&1
# (4) the new kernels: gTopk tree, TopkA2 re-selection, gradient landing
$NCU --kernel-name regex:gtopk_kernel --launch-skip 4 -f -o gpurun_out/gtopk_vgg python -m oktopk_b200.bench.sweep --schemes gtopk --iters 4 --warmup 2 --sizes 14728266 --densities 0.001 > gpurun_out/ncu_gtopk.log 2>&1
$NCU --kernel-name regex:land_kernel --launch-skip 4 -f -o gpurun_out/land python bench.py --steps 3 --warmup 4 --dense-warmup 0 --no-e2e --no-extra --no-graph > gpurun_out/ncu_land.log 2>&1
for r in okt_vgg_d001 okt_128m_d001 okt_128m_d01 gtopk_vgg land; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
ncu -i gpurun_out/okt_vgg_d001.ncu-rep --page source --csv --print-source sass > gpurun_out/okt_vgg_d001.source.csv 2>/dev/null
rm -f gpurun_out/okt_128m_d001.ncu-rep gpurun_out/okt_128m_d01.ncu-rep gpurun_out/land.ncu-rep gpurun_out/gtopk_vgg.ncu-rep
ls -la gpurun_out/*.csv gpurun_out/*.ncu-rep
