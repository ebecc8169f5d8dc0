#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/task_bench.py 128 128 128 3 2>&1 | tail -1
python scripts/task_bench.py 256 256 256 3 2>&1 | tail -1
