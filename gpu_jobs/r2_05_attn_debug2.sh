#!/bin/bash
# round 2, call 5: is the tmem attention failure order / PDL dependent?
echo "=== alone"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_attention_tmem and 257" 2>&1 | tail -3
echo "=== alone batch-invariant"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "batch_invariant" 2>&1 | tail -3
echo "=== all attention, PDL default"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -4
echo "=== all attention, PDL=0"; CNB_PDL=0 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -4
echo "=== all attention, PDL=1"; CNB_PDL=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -4
echo "=== tmem tests only"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention_tmem" 2>&1 | tail -4
