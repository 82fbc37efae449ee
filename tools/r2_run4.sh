#!/bin/bash
# final sanity: the whole GPU suite (hard-bounded) + smoke
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_tests.txt; cat gpurun_out/r2_tests.txt
timeout -k 5 200 python __graft_entry__.py smoke 2>&1 | tail -3
