#!/bin/bash
# torchrun equivalent of the reference job script (VGG/vgg16_oktopk.sh)
exec "$(dirname "$0")/run.sh" vgg16 oktopk "${NGPUS:-8}" "$@"
