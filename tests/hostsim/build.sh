#!/bin/sh
# builds the CPU simulation of the kernels' rule code (test infrastructure)
set -e
cd "$(dirname "$0")"
mkdir -p _build
g++ -O2 -std=c++17 -shared -fPIC -DQK_HOSTSIM_PLAYOUT -I ../../quantik-core-py_b200/csrc -o _build/libqk_hostsim.so hostsim.cpp
