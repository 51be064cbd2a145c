#!/bin/bash
# usage: tools/sass.sh <kernel-name-regex>  -> /tmp/k.sass (one instruction per line)
cuobjdump -sass /root/repo/doper_b200/_lib/libdoper_b200.so | awk -v pat="$1" '/Function : /{p=($0 ~ pat)} p{print}' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' | sed 's/ *$//' > /tmp/k.sass
wc -l /tmp/k.sass
