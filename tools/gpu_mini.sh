#!/bin/bash
# last seconds of the round's GPU budget: smoke + the order-2 search parity tests of the committed build
timeout 24 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 20 -k "order2_matches_oracle or smoke" 2>&1 | tail -3
