#!/bin/bash
# usage: tools/sass_fn.sh <substring of the mangled kernel name> [lib]  -> SASS of that kernel on stdout
lib=${2:-kamino_b200/_lib/libkamino_b200.so}
cuobjdump -sass "$lib" | awk -v pat="$1" '/Function :/ {on = index($0, pat) > 0} on {print}'
