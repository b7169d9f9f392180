#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q -rP 2>&1 | grep -E "passed|failed|rel-L2|SNR|Error|error|assert" | head -30
