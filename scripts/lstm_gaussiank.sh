#!/bin/bash
# torchrun equivalent of the reference job script (LSTM/lstm_gaussiank.sh)
exec "$(dirname "$0")/run.sh" lstman4 gaussiank "${NGPUS:-8}" "$@"
