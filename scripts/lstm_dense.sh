#!/bin/bash
# torchrun equivalent of the reference job script (LSTM/lstm_dense.sh)
exec "$(dirname "$0")/run.sh" lstman4 dense "${NGPUS:-8}" "$@"
