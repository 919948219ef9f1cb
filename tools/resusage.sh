#!/bin/bash
# usage: tools/resusage.sh file.o pattern  -- registers / stack of the kernels whose demangled name matches the pattern
cuobjdump -res-usage "$1" 2>/dev/null | c++filt | awk '/Function/ {name=$0} /REG:/ {print name " | " $0}' | grep "$2" | sed 's/ Function void splacu:://; s/(splacu::Semiring[^|]*|/ |/; s/SHARED.*//'
