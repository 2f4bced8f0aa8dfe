"""Build a HOST-executable copy of a CUDA source of the product (test infrastructure; see cuda_emu.h).

  python tests/cuda_emu/build_emu.py            -> tests/cuda_emu/_build/libnn_backward_emu.so

The .cu file is used as it is, with three mechanical rewrites: `kernel<<<grid, block, smem, stream>>>(args)` becomes
EMU_LAUNCH(kernel, grid, block, smem, stream, args); `extern __shared__ T name[]` (dynamic shared memory) becomes a
file-scope array; the relative include of the C-ABI header is made absolute.  Then g++ -std=c++20 compiles it against
the emulation headers in this directory (which shadow <cuda_runtime.h> / <cuda_bf16.h>)."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "guidedvd-3dgs_b200", "csrc")
OUT = os.path.join(HERE, "_build")

LAUNCH = re.compile(r"(\w+)<<<(.*?)>>>\s*\(", re.S)
DYN_SHARED = re.compile(r"extern\s+__shared__\s+((?:__align__\(\d+\)\s+)?)([\w ]+?)\s+(\w+)\[\];")


def rewrite(src):
    out, pos = [], 0
    for m in LAUNCH.finditer(src):
        # find the matching ')' of the argument list
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args = src[m.end():i - 1]
        out.append(src[pos:m.start()])
        out.append(f"EMU_LAUNCH({m.group(1)}, {m.group(2)}, {args})")
        pos = i
    out.append(src[pos:])
    src = "".join(out)
    dyn = []

    def hoist(m):
        dyn.append(f"alignas(16) {m.group(2)} {m.group(3)}[1 << 17];")
        return f"/* dynamic shared memory: file-scope array {m.group(3)} */"
    src = DYN_SHARED.sub(hoist, src)
    src = re.sub(r'#include "\.\./\.\./include/(\w+\.h)"', lambda m: f'#include "{os.path.join(ROOT, "include", m.group(1))}"', src)
    # the file-scope arrays must be visible before their first use: put them right after the includes
    block = "namespace {\n" + "\n".join(dyn) + "\n}\n" if dyn else ""
    marker = "extern thread_local std::string g_nn_err_ext;"  # defined in gemm_tc.cu, which is not part of a host build
    if marker in src:
        src = src.replace(marker, "thread_local std::string g_nn_err_ext;\n" + block, 1)
    elif block:
        last = max(m.end() for m in re.finditer(r"^#include .*$", src, re.M))
        src = src[:last] + "\n" + block + src[last:]
    return src


def build(name="nn_backward"):
    os.makedirs(OUT, exist_ok=True)
    cu = os.path.join(CSRC, name + ".cu")
    cpp = os.path.join(OUT, name + "_emu.cpp")
    so = os.path.join(OUT, f"lib{name}_emu.so")
    if os.path.exists(so) and os.path.getmtime(so) > max(os.path.getmtime(cu), os.path.getmtime(os.path.join(HERE, "cuda_emu.h")),
                                                        os.path.getmtime(__file__)):
        return so
    with open(cu) as f:
        text = rewrite(f.read())
    with open(cpp, "w") as f:
        f.write(text)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", HERE, "-o", so, cpp])
    return so


if __name__ == "__main__":
    print(build(*sys.argv[1:]))
