#!/usr/bin/env python
"""Write the C code block of INTEGRATION.md section 2 ("The shim a maintainer adds") to a file, verbatim,
so that tests/shim_c99_main.c compiles exactly what the document shows."""
import re
import sys

text = open(sys.argv[1]).read()
sec = text[text.index("## 2."):]
m = re.search(r"```c\n(.*?)```", sec, re.S)
open(sys.argv[2], "w").write("/* extracted verbatim from INTEGRATION.md section 2 by tools/extract_shim.py */\n" + m.group(1))
