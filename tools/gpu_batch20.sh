#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest -q -x -s -m gpu tests/test_frontend.py -k "weibel_growth" 2>&1 | grep -E "^weibel|^E  |assert|Error" | cut -c1-900 | head -20
