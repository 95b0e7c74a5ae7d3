#!/bin/bash
# ncu captures of the hot kernels on ONE GPU (B200_PROFILING.md recipe: --set full --clock-control none, one kernel per
# capture, after warm-up).  Reports + csv exports land in gpurun_out/; summaries are distilled into profiles/ by
# scripts/ncu_summary.py.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none --launch-count 1"
SW="python -m oktopk_b200.bench.sweep --schemes oktopk --iters 6 --warmup 3"
# steady-state (threshold-reuse) Ok-Topk call, VGG-16-sized bucket, density 0.001
$NCU --kernel-name regex:oktopk_fused --launch-skip 6 -f -o gpurun_out/okt_vgg_d001 $SW --sizes 14728266 --densities 0.001 > gpurun_out/ncu_okt_vgg.log 2>&1
# BERT-sized bucket (128 Mi), density 0.001 and 0.1
$NCU --kernel-name regex:oktopk_fused --launch-skip 6 -f -o gpurun_out/okt_128m_d001 $SW --sizes 128M --densities 0.001 > gpurun_out/ncu_okt_128m_d001.log 2>&1
$NCU --kernel-name regex:oktopk_fused --launch-skip 6 -f -o gpurun_out/okt_128m_d1 $SW --sizes 128M --densities 0.1 > gpurun_out/ncu_okt_128m_d1.log 2>&1
# exact-threshold iteration (the 1-in-32 flavour): first call of a fresh engine
$NCU --kernel-name regex:oktopk_fused --launch-skip 0 -f -o gpurun_out/okt_128m_exact $SW --sizes 128M --densities 0.001 > gpurun_out/ncu_okt_exact.log 2>&1
# fused optimizer updates
ncu --set full --clock-control none --launch-count 1 --kernel-name regex:fused_sgd --launch-skip 20 -f -o gpurun_out/fused_sgd python bench.py --no-graph --no-e2e --steps 5 --warmup 30 --dense-warmup 0 > gpurun_out/ncu_sgd.log 2>&1
ncu --set full --clock-control none --launch-count 1 --kernel-name regex:fused_bert_adam --launch-skip 6 -f -o gpurun_out/fused_adam python bench.py --model bert --no-graph --no-e2e --steps 3 --warmup 4 > gpurun_out/ncu_adam.log 2>&1
for r in okt_vgg_d001 okt_128m_d001 okt_128m_d1 okt_128m_exact fused_sgd fused_adam; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
for r in okt_vgg_d001 okt_128m_d1; do
  ncu -i gpurun_out/$r.ncu-rep --page source --csv --print-source sass > gpurun_out/$r.source.csv 2>/dev/null
done
rm -f gpurun_out/okt_128m_d001.ncu-rep gpurun_out/okt_128m_exact.ncu-rep gpurun_out/fused_adam.ncu-rep
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
