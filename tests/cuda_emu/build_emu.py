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

LAUNCH = re.compile(r"([\w:]+(?:<[^<>;()]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)  # kernel or kernel<targs>
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
        out.append(f"EMU_LAUNCH(({m.group(1)}), {m.group(2)}, {args})")  # parenthesised: template arguments may hold commas
        pos = i
    out.append(src[pos:])
    src = "".join(out)
    dyn = []

    def hoist(m):
        # one array per NAME (kernels of a file may share it: blocks never overlap); 256 KB-aligned so that the low 18 bits
        # of a shared-memory pointer are its offset in the block's window (tc_emu.h::smem_u32, descriptor start addresses)
        decl = f"alignas(262144) {m.group(2)} {m.group(3)}[1 << 18];"
        if decl not in dyn:
            dyn.append(decl)
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


def build(name="nn_backward", sources=None, extra_flags=()):
    """name: a single csrc/<name>.cu, or the library name when `sources` lists several files that link into one .so."""
    os.makedirs(OUT, exist_ok=True)
    sources = list(sources or [name + ".cu"])
    asan = os.environ.get("GVD_EMU_ASAN") == "1"   # tools/emu_memcheck.sh: AddressSanitizer build of the same sources
    if asan:
        extra_flags = tuple(extra_flags) + ("-fsanitize=address", "-fno-omit-frame-pointer")
    so = os.path.join(OUT, f"lib{name}_emu{'_asan' if asan else ''}.so")
    deps = [os.path.join(CSRC, s) for s in sources] + [os.path.join(HERE, h) for h in ("cuda_emu.h", "tc_emu.h", "cuda.h", "cudaTypedefs.h")]
    deps += [os.path.join(HERE, "cub", "cub.cuh"), __file__]
    deps += [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")]
    if os.path.exists(so) and os.path.getmtime(so) > max(os.path.getmtime(d) for d in deps):
        return so
    cpps = []
    for src in sources:
        with open(os.path.join(CSRC, src)) as f:
            text = rewrite(f.read())
        cpp = os.path.join(OUT, os.path.splitext(src)[0] + ("_emu_asan.cpp" if asan else "_emu.cpp"))
        with open(cpp, "w") as f:
            f.write(text)
        cpps.append(cpp)
    # -I CSRC: the rewritten copies live in _build/ but still include their neighbours ("raster_common.cuh")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-DGVD_HOST_EMU", "-D__CUDACC__",
                           "-x", "c++", "-I", HERE, "-I", CSRC, *extra_flags, "-o", so, *cpps])
    return so


RASTER_SOURCES = ["raster_api.cu", "raster_forward.cu", "raster_backward.cu", "grad_exchange.cu"]


if __name__ == "__main__":
    if sys.argv[1:] == ["raster"]:
        print(build("raster", RASTER_SOURCES))
    else:
        print(build(*sys.argv[1:]))
