#!/bin/bash
# usage: scripts/kinfo.sh <NX> — resource usage of the cached kernel with that NX, and where its pivot loop is
f=$(grep -l "^#define NX $1\$" mcp_b200/_kcache/*.cu | xargs ls -t | head -1)
b=${f%.cu}.cubin
echo $f
cuobjdump -res-usage $b 2>/dev/null | grep -A1 "Function mcp_" | grep -o "Function mcp_[a-z_0-9]*\|REG:[0-9]*\|STACK:[0-9]*" | paste - - -
