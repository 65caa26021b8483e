#!/bin/bash
# SASS listings of the hot kernels for profiles/ (cuobjdump reads the built .so here; no GPU needed).
# usage: scripts/dump_sass.sh [tag]
TAG=${1:-r02}
LIB=raynet_b200/libraynet_b200.so
dump() {  # name mangled
  {
    echo "# cuobjdump -sass -fun $2 $LIB   ($(date -u +%F))"
    cuobjdump -res-usage -fun "$2" $LIB 2>/dev/null | grep -E "REG:|Function"
    echo "# mnemonic histogram (memory / async / atomic instructions first)"
    cuobjdump -sass -fun "$2" $LIB 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/^@[!UP0-9T]* //' \
      | sort | uniq -c | sort -rn | awk '{printf "%6d %s\n", $1, $2}' | grep -E "LDG|STG|RED|ATOM|LDGSTS|UBLKCP|UTMA|UTC|LDTM|LDS|STS|SYNCS|MUFU|SHFL|BAR|FFMA2|FADD2|DADD|DFMA|DMUL" | head -40
    echo "# full listing"
    cuobjdump -sass -fun "$2" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/[[:space:]]+$//'
  } > profiles/${TAG}_sass_$1.txt
  echo "profiles/${TAG}_sass_$1.txt: $(wc -l < profiles/${TAG}_sass_$1.txt) lines"
}
dump bp4_nch4_next _Z10bp4_kernelILi4ELb0EEv5RnDev7Bp2Args
dump bp4_nch3_first _Z10bp4_kernelILi3ELb1EEv5RnDev7Bp2Args
dump simscore3_v9 _Z16simscore3_kernelILi9ELi16EEv5RnDev10SimMapArgs
dump planemap3 _Z16planemap3_kernel5RnDev10SimMapArgs
dump depth3 _Z13depth3_kernel5RnDev10Depth2Args
dump peer_allreduce _Z21peer_allreduce_kernelILi8EEv8PeerArgs
dump peer_allreduce_mc _Z24peer_allreduce_mc_kernel8PeerArgs
dump bp_parity _Z16bp_parity_kernelILb0EEv5RnDev10ParityArgs
dump bp4_first_mapped_nch4 _Z23bp4_first_mapped_kernelILi4EEv5RnDev9FirstArgs
dump conv3x3_tc _Z17conv3x3_tc_kernel14CUtensorMap_stS_10ConvTcArgs
