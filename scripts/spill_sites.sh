#!/bin/bash
# Where does a variant spill?  usage: scripts/spill_sites.sh ILi2ELb0ELb1ELb0ELb0E   (mangled template args; needs /tmp/libcpbus_ptxas.so from ptxas_summary.sh)
cuobjdump -sass /tmp/libcpbus_ptxas.so | awk -v pat="$1" '/Function : /{f=($3 ~ pat)} f' | grep -v '^\s*/\* 0x' | grep -v '^\s*$' > /tmp/variant.sass
grep -n "STL\|LDL\|STG.E.ENL2.256\|BAR.SYNC\|ACQBULK\|PREEXIT\|BRA" /tmp/variant.sass | awk '{print $1, $2, $3, $4, $5, $6}' | grep -B2 -A2 "STL\|LDL"
