#!/bin/bash
# torchrun equivalent of the reference job script (VGG/vgg16_dense.sh)
exec "$(dirname "$0")/run.sh" vgg16 dense "${NGPUS:-8}" "$@"
