#!/bin/bash
mkdir -p gpurun_out
IGV_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -s -k "frame_step_graph_replay and 1" > gpurun_out/dbg_pytest.log 2>&1
grep -E "igv_frame_step|Error|passed|failed" gpurun_out/dbg_pytest.log | head -40
