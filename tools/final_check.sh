#!/bin/bash
timeout 900 python -m pytest tests/test_headline_geometry_gpu.py tests/test_ctvit_gpu.py -q -s 2>&1 | grep -E "grid 32|median|passed|failed|Error|assert" | cut -c1-600
