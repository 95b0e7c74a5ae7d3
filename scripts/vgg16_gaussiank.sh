#!/bin/bash
# torchrun equivalent of the reference job script (VGG/vgg16_gaussiank.sh)
exec "$(dirname "$0")/run.sh" vgg16 gaussiank "${NGPUS:-8}" "$@"
