#!/bin/bash
# torchrun equivalent of the reference job script (BERT/bert/bert_dense.sh)
exec "$(dirname "$0")/run.sh" bert dense "${NGPUS:-8}" "$@"
