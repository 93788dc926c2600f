#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/debug_vgg2.py > $O/f_debug_vgg2.log 2>&1; tail -30 $O/f_debug_vgg2.log
timeout 120 tools/micro/xfer_probe > $O/f_xfer_probe.log 2>&1; cat $O/f_xfer_probe.log
timeout 900 python tools/check_bwd_protocols.py > $O/f_check_bwd.log 2>&1; cat $O/f_check_bwd.log
