#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_ingest_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/quick_fan.py 400 8 3 2>&1 | grep -v Voxel
timeout 300 python tools/quick_fan.py 400 8 3 u16 2>&1 | grep -v Voxel
