#!/bin/bash
# torchrun equivalent of the reference job script (VGG/vgg16_topkA.sh)
exec "$(dirname "$0")/run.sh" vgg16 topkA "${NGPUS:-8}" "$@"
