#!/bin/bash
# torchrun equivalent of the reference job script (LSTM/lstm_gtopk.sh)
exec "$(dirname "$0")/run.sh" lstman4 gtopk "${NGPUS:-8}" "$@"
