#!/bin/bash
# Regenerates pokegym_b200/csrc/gb_classes.inc from the per-opcode base table of gb_predecode.h (run after changing pd_build_base).
cd "$(dirname "$0")/.."
g++ -std=c++17 -I pokegym_b200/csrc -o /tmp/gen_classes tools/gen_classes.cpp && /tmp/gen_classes > pokegym_b200/csrc/gb_classes.inc.new && mv pokegym_b200/csrc/gb_classes.inc.new pokegym_b200/csrc/gb_classes.inc && tail -1 pokegym_b200/csrc/gb_classes.inc
