#!/bin/bash
# Run the six schemes of one workload back to back (reference: */sbatch_*_jobs.sh submit them to SLURM).
w="${1:?workload: vgg16|lstm|bert}"; shift
for s in dense topkA topkDSA gtopk gaussiank oktopk; do "$(dirname "$0")/${w}_${s}.sh" "$@"; done
