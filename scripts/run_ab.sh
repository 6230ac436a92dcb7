#!/bin/bash
# A/B of two kernel-variant libraries built by scripts/kbench.py (development tool): scripts/run_ab.sh '<json variants>' <N list>
export KBENCH_VARIANTS="$1"
export KBENCH_N="$2"
python scripts/kbench.py run
