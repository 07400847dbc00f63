#!/bin/bash
O=gpurun_out/host; mkdir -p $O
timeout 300 python tools/host_overhead.py > $O/host.log 2>&1; echo "rc=$?" >> $O/host.log
