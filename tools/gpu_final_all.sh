#!/bin/bash
# last call of the round: the whole GPU suite, then both bench arms and the two kernel breakdowns
./tools/gpu_full_suite.sh
./tools/gpu_round_final_light.sh ${1:-r02h}
