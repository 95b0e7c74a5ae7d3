#!/bin/bash
# torchrun equivalent of the reference job script (VGG/vgg16_topkDSA.sh)
exec "$(dirname "$0")/run.sh" vgg16 topkDSA "${NGPUS:-8}" "$@"
