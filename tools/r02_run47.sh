#!/bin/bash
# round-2 GPU call 47: attention / forward tests after the slot-counter fix
python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -q 2>&1 | tail -3
