#!/bin/bash
for pe in 0 1 2 3; do GVD_FLASH=v7 GVD_FLASH_POLY=$pe timeout 200 python tools/bench_attn.py 2>&1 | tail -3 | sed "s/^/poly=$pe /"; done
