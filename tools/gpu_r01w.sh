#!/bin/bash
mkdir -p gpurun_out
PYTHONPATH=. python tools/subrange_timing.py rootfilter
