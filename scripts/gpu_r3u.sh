#!/bin/bash
set -x
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
