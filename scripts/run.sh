#!/bin/bash
# One launcher for every (workload x scheme) job of the reference (VGG/vgg16_*.sh, LSTM/lstm_*.sh, BERT/bert/bert_*.sh:
# 3 workloads x {oktopk, topkA, topkDSA, gtopk, gaussiank, dense}).  SLURM `sbatch` + `srun python -m mpi4py` becomes
# torchrun with one process per GPU on one node; the sparse allreduce runs on the fused peer-memory kernels.
#
#   scripts/run.sh <workload: vgg16|lstman4|bert|resnet20|resnet50|alexnet|lstm> <scheme> [ngpus] [extra cli flags...]
#   density=0.001 scripts/run.sh vgg16 oktopk 8 --max-iters 200
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
workload="${1:?workload}"; scheme="${2:?scheme: oktopk|topkA|topkDSA|gtopk|gaussiank|dense|...}"; ngpus="${3:-1}"
shift $(( $# < 3 ? $# : 3 ))
extra=""
source "$here/exp_configs/$workload.conf"
case "$scheme" in
  dense|none) comp="--compressor none" ;;
  topkDSA)    comp="--compression --compressor topkSA" ;;
  *)          comp="--compression --compressor $scheme" ;;
esac
cd "$here/.."
cmd=(python -m torch.distributed.run --nnodes=1 --nproc-per-node "$ngpus" --master-addr 127.0.0.1
     --master-port "${MASTER_PORT:-29531}" -m oktopk_b200.train.cli
     --dnn "$dnn" --dataset "$dataset" --preset "$preset" --lr "$lr" --batch-size "$batch_size"
     --max-epochs "$max_epochs" --nsteps-update "$nstepsupdate" --density "$density" --nworkers "$ngpus" $comp $extra)
[ -n "${data_dir:-}" ] && cmd+=(--data-dir "$data_dir")
echo "${cmd[@]}" "$@"
exec "${cmd[@]}" "$@"
