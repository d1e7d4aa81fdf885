#!/bin/bash
# scratch: dump one kernel's SASS to /tmp/sass/<tag>.sass.  usage: tools/sass_loop.sh <substring of mangled name> <tag> [lib]
lib=${3:-roomslam_b200/libroomslam_b200.so}
mkdir -p /tmp/sass
f=$(cuobjdump -sass $lib | grep "Function :" | grep "$1" | head -1 | awk '{print $3}')
cuobjdump -sass -fun "$f" $lib 2>/dev/null | grep -v "^\s*/\* 0x" | grep "^\s*/\*[0-9a-f]*\*/" > /tmp/sass/$2.sass
echo "$f: $(wc -l < /tmp/sass/$2.sass) instructions"
