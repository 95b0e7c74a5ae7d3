#!/bin/bash
# torchrun equivalent of the reference job script (LSTM/lstm_oktopk.sh)
exec "$(dirname "$0")/run.sh" lstman4 oktopk "${NGPUS:-8}" "$@"
