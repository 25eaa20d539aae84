#!/usr/bin/env python
"""Print registers / spills per kernel from a `build.py --ptxas` log: tools/regs.py /tmp/build.log [filter]"""
import re, subprocess, sys
log = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ents = re.findall(r"Compiling entry function '(\w+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log)
for name, stack, ss, sl, regs in ents:
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'\(.*', '', dem)
    if flt in dem:
        print(f"{regs:>4} regs stack={stack:>5} spill={ss}/{sl}  {dem[:150]}")
