#!/bin/bash
set -u
O=gpurun_out/tc7; mkdir -p $O
timeout 300 python tools/tc_scorer_check.py 4096 > $O/check.log 2>&1; rc=$?; echo "rc=$rc" >> $O/check.log
