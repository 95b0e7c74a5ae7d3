#!/bin/bash
# torchrun equivalent of the reference job script (BERT/bert/bert_gaussiank.sh)
exec "$(dirname "$0")/run.sh" bert gaussiank "${NGPUS:-8}" "$@"
