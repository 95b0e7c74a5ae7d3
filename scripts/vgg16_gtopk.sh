#!/bin/bash
# torchrun equivalent of the reference job script (VGG/vgg16_gtopk.sh)
exec "$(dirname "$0")/run.sh" vgg16 gtopk "${NGPUS:-8}" "$@"
