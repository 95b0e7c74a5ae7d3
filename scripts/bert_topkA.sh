#!/bin/bash
# torchrun equivalent of the reference job script (BERT/bert/bert_topkA.sh)
exec "$(dirname "$0")/run.sh" bert topkA "${NGPUS:-8}" "$@"
