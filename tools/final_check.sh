#!/bin/bash
timeout 900 python -m pytest tests/test_ctclip_gpu.py -q -s 2>&1 | grep -E "median|passed|failed|Error|assert" | cut -c1-300
