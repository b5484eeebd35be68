#!/bin/bash
set -u
timeout 300 python tools/quick_e2e.py 400 1 0 2>/dev/null
timeout 300 python tools/quick_e2e.py 400 8 3 2>/dev/null
