#!/bin/bash
# parity of the pixel pass + a timing sweep
python -m pytest tests -m gpu -x -q -k "parity or pixel or batch or multi or golden" 2>&1 | grep -E "Error|assert|passed|failed" | cut -c1-400 | tail -12
bash ${1:-tools/gpu_sweep2.sh} 2>&1 | grep fps
