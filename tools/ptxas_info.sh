#!/bin/bash
# scratch: registers / spills per kernel of one translation unit.  usage: tools/ptxas_info.sh rec_pair.cu [extra nvcc flags]
f=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I include -Xptxas=-v "$@" \
  -c roomslam_b200/csrc/$f -o /tmp/ptxas_info.o 2>&1 | awk '
  /Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)}
  /spill stores/ {spill=$0; sub(/^ */,"",spill)}
  /Used [0-9]+ registers/ {regs=$0; sub(/.*Used /,"",regs); sub(/ registers.*/,"",regs); print regs " regs | " spill " | " name}' | c++filt | sed 's/(anonymous namespace):://; s/((anonymous namespace)::[A-Za-z]*)//'
