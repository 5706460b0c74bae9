#!/usr/bin/env python3
"""Make an UNMODIFIED reference problem file (problems/<name>.cpp, the link-time plugin of makefile:22,28-29) compilable
for the device: writes a copy in which

  * every free function defined at file scope gets ``__host__ __device__`` (the hooks call such helpers, e.g. dusty_wind's
    vapour-pressure functions), and
  * every non-const file-scope ``static`` scalar becomes ``__managed__`` (hooks keep state there, e.g. ``T_eq``, which
    user_initial_conditions() sets on the host and user_boundary_left() reads inside the hydro kernel).

The five ``c_Species::user_*`` members get their execution space from the facade header (aiolos_b200/plugin/aiolos.h) and are
left untouched.  Everything else is copied verbatim, line for line (a ``#line`` directive keeps compiler messages pointing
at the original file).  usage: prepare_problem.py <in.cpp> <out.cu>"""
import re
import sys

FUNC = re.compile(r"^\s*(static\s+|inline\s+)*[A-Za-z_][\w:<>\s\*&]*?\b([A-Za-z_]\w*)\s*\([^;{}]*\)\s*(const\s*)?(\{.*)?$")
VAR = re.compile(r"^(\s*)static\s+(?!const\b)((?:unsigned\s+|long\s+)*(?:double|float|int|long|bool|char))\s+([A-Za-z_]\w*\s*(=[^;]*)?;)")
SKIP = ("#", "//", "namespace", "struct", "class", "typedef", "using", "template", "extern", "enum", "}", "return")


def transform(src: str, name: str) -> str:
    out = ['#line 1 "%s"' % name]
    depth = 0
    in_block_comment = False
    for line in src.splitlines():
        stripped = line.strip()
        new = line
        if depth == 0 and not in_block_comment and stripped and not stripped.startswith(SKIP):
            m = VAR.match(line)
            if m:
                new = f"{m.group(1)}static __managed__ {m.group(2)} {m.group(3)}" + line[m.end():]
            else:
                f = FUNC.match(line)
                if f and "::" not in line.split("(")[0]:
                    new = "__host__ __device__ " + line.lstrip()
        out.append(new)
        # brace depth, ignoring comments and string/char literals
        code = re.sub(r'"(\\.|[^"\\])*"|\'(\\.|[^\'\\])*\'', "", line)
        if in_block_comment:
            end = code.find("*/")
            if end < 0:
                continue
            code = code[end + 2:]
            in_block_comment = False
        code = re.sub(r"/\*.*?\*/", "", code)
        start = code.find("/*")
        if start >= 0:
            code = code[:start]
            in_block_comment = True
        code = code.split("//")[0]
        depth += code.count("{") - code.count("}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    with open(sys.argv[1]) as f:
        text = f.read()
    with open(sys.argv[2], "w") as f:
        f.write(transform(text, sys.argv[1]))
