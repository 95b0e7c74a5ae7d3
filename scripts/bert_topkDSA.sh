#!/bin/bash
# torchrun equivalent of the reference job script (BERT/bert/bert_topkDSA.sh)
exec "$(dirname "$0")/run.sh" bert topkDSA "${NGPUS:-8}" "$@"
