#!/bin/bash
# torchrun equivalent of the reference job script (LSTM/lstm_topkDSA.sh)
exec "$(dirname "$0")/run.sh" lstman4 topkDSA "${NGPUS:-8}" "$@"
