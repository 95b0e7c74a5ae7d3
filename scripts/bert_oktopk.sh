#!/bin/bash
# torchrun equivalent of the reference job script (BERT/bert/bert_oktopk.sh)
exec "$(dirname "$0")/run.sh" bert oktopk "${NGPUS:-8}" "$@"
