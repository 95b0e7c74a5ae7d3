#!/bin/bash
# torchrun equivalent of the reference job script (LSTM/lstm_topkA.sh)
exec "$(dirname "$0")/run.sh" lstman4 topkA "${NGPUS:-8}" "$@"
