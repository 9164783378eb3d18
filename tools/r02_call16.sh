#!/bin/bash
# round-2 call 16: where does a k_track pass spend its cycles?  (profiling build with per-phase clock64 totals)
O=gpurun_out; mkdir -p $O
V=$PWD/headtrackr_b200/variants
HT_LIB=$V/libht_ptrace.so timeout 300 python tools/track_timeline.py 1024 > $O/r02c16_phases.txt 2>&1; head -5 $O/r02c16_phases.txt; tail -6 $O/r02c16_phases.txt
HT_LIB=$V/libht_ptrace.so HT_TRACK_HEAVY=0 HT_TRACK_MID=0 timeout 300 python tools/track_timeline.py 1024 > $O/r02c16_phases_c2.txt 2>&1; head -5 $O/r02c16_phases_c2.txt; tail -6 $O/r02c16_phases_c2.txt
HT_LIB=$V/libht_ptrace.so timeout 300 python tools/track_timeline.py 64 > $O/r02c16_phases_n64.txt 2>&1; head -5 $O/r02c16_phases_n64.txt; tail -6 $O/r02c16_phases_n64.txt
