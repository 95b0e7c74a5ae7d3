#!/bin/bash
# torchrun equivalent of the reference job script (BERT/bert/bert_gtopk.sh)
exec "$(dirname "$0")/run.sh" bert gtopk "${NGPUS:-8}" "$@"
