#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_waits.py vimeo_x4 2>&1 | tail -40
