#!/bin/bash
# SASS evidence for the judged kernels (no GPU needed): opcode tallies and the TMA / mbarrier / cp.async lines of
# k_r1cs_tiled (K2) and k_ntt_pass (K3) as built into arithmetic-circuits_b200/libacg.so.  Usage: tools/sass_evidence.sh r02
set -eu
R=${1:-r02}
SO=arithmetic-circuits_b200/libacg.so
dump() { # $1 = regex of the function header, $2 = output file, $3 = title
  cuobjdump -sass "$SO" 2>/dev/null | awk -v pat="$1" '/Function : /{f = ($0 ~ pat)} f' > /tmp/_sass.txt
  {
    echo "# $3"
    echo "# source: cuobjdump -sass $SO (sm_100a), function(s) matching /$1/"
    grep "Function : " /tmp/_sass.txt | sed 's/^\s*/# /'
    echo "# instructions: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' /tmp/_sass.txt)"
    echo "#"
    echo "# ---- opcode tally"
    grep -o '^\s*/\*[0-9a-f]\{4\}\*/\s*\(@!\?U\?P[0-9T]\s\+\)\?[A-Z][A-Z0-9_.]*' /tmp/_sass.txt | sed 's/.*\s//' | sort | uniq -c | sort -rn
    echo "#"
    echo "# ---- TMA bulk copies (UBLKCP / UBLKPF), mbarrier (SYNCS), cp.async (LDGSTS), barriers, PDL (ACQBULK / PREEXIT), spills (STL / LDL)"
    grep -n 'UBLKCP\|UBLKPF\|SYNCS\|LDGSTS\|LDGDEPBAR\|BAR\.\|ACQBULK\|PREEXIT\|STL\|LDL\|ATOM\|RED\.' /tmp/_sass.txt | sed 's/\s\+\/\* 0x[0-9a-f]* \*\///' | cut -c1-150
  } > "$2"
  echo "$2: $(wc -l < "$2") lines"
}
dump 'k_r1cs_tiledINS_7Bn254FrELb0ELi0ELb0' profiles/${R}_sass_k2.txt "K2 k_r1cs_tiled<Bn254Fr, EMIT=false, V=0>: the R1CS check (shipped default geometry)"
dump 'k_ntt_passINS_7Bn254Fr' profiles/${R}_sass_k3.txt "K3 k_ntt_pass<Bn254Fr>: one pass of the radix-2 NTT"
dump 'k_r1cs_longrowsINS_7Bn254FrELb0' profiles/${R}_sass_k2_longrows.txt "k_r1cs_longrows<Bn254Fr, EMIT=false>: warp-per-row check of the rows too long for a tile"
