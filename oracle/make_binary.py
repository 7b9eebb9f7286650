"""Builds the REAL `muchsalsa` executable with the GPU path wired in — TEST INFRASTRUCTURE, build container only.

    python oracle/make_binary.py            ->  oracle/_ref/muchsalsa   (git-ignored; travels to the GPU box)

What a maintainer would do by hand, done by script so that it can be re-run and checked:
  1. the reference's `src/*.cpp` and `libms/src/kernel/cc.cpp` are COPIED to a scratch directory (default
     /tmp/msbin_build; nothing of the reference enters the repository),
  2. the eight `T const *const` container element types libstdc++ rejects lose their inner const (the reference is
     built with libc++, which accepts them: src/main.cpp:211,212,247,486,495,541,597, kernel/cc.cpp:46; the
     `std::any_cast<Edge const *const>` of main.cpp:418 is legal and stays),
  3. main() gets the INTEGRATION.md §2 change (and, for the sequence fetches, the §6 change: MUCHSALSA_GPU_SEQ=1 loads
     the two sequence files into the resident store and runs every assemblePath call as record -> one batched fetch ->
     replay; the two getters of a COPY of SequenceUtils.cpp forward to the active batch — oracle/binary_patch/*.inc): with MUCHSALSA_GPU=<device> in the environment the two statements
     `blastFileReader.read(); matchMap.calculateEdges();` (src/main.cpp:156-157) are replaced by
     `muchsalsa::gpu::constructGraph(...)`; with MUCHSALSA_GPU_STAGES=1 in addition the job loops over
     chainingAndOverlaps / findContractionEdges / findDeletableEdges / computeBitweight (src/main.cpp:170-191,
     247-256, 263-270) become the shim's one-call forms.  Without MUCHSALSA_GPU the binary runs the reference's own
     code, statement for statement,
  4. everything is compiled with g++ -std=c++20 -O2 against the reference's headers and linked with the reference's
     objects (oracle/_ref/obj, built by oracle/Makefile), the host shim and libmsgpu.so.
tests/test_binary_gpu.py runs the result on config 1 with and without MUCHSALSA_GPU and compares the three output
files byte for byte.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(ROOT, "muchsalsa_b200")

GPU_CONSTRUCT = '''    // Read BLAST file + calculate edges: on the GPU when MUCHSALSA_GPU=<device> is set, else the CPU path as before
    std::unique_ptr<muchsalsa::gpu::Context> pGpuCtx;
    bool                                     gpuStages = false;
    if (char const *const gpuEnv = std::getenv("MUCHSALSA_GPU")) {
      pGpuCtx   = std::make_unique<muchsalsa::gpu::Context>(std::atoi(gpuEnv));
      gpuStages = std::getenv("MUCHSALSA_GPU_STAGES") != nullptr;
      muchsalsa::gpu::constructGraph(*pGpuCtx, app.getContigsFilePath(), graph, matchMap, registryNanopore,
                                     registryIllumina);
    } else {
      auto bfAccessor = BlastFileAccessor(app.getContigsFilePath());
      auto blastFileReader =
          BlastFileReader(&threadPool, &bfAccessor, &graph, &matchMap, &registryNanopore, &registryIllumina);
      blastFileReader.read();
      matchMap.calculateEdges();
    }
'''


def patch_main(src: str) -> str:
    # 2. libstdc++: no const element types in node containers
    src, n = re.subn(r"(std::set<(?:Vertex|Edge) const \*)const>", r"\1>", src)
    assert n == 7, f"expected 7 std::set element types to fix, found {n}"
    # 3a. construction (src/main.cpp:152-157)
    old = '''    // Read BLAST file
    auto bfAccessor = BlastFileAccessor(app.getContigsFilePath());
    auto blastFileReader =
        BlastFileReader(&threadPool, &bfAccessor, &graph, &matchMap, &registryNanopore, &registryIllumina);
    blastFileReader.read();
    matchMap.calculateEdges();
'''
    assert old in src, "the construction statements of main() have moved"
    src = src.replace(old, GPU_CONSTRUCT, 1)
    src = src.replace('#include "Application.h"', '#include <cstdlib>\n\n#include "Application.h"\n#include "ms_gpu_shim.h"', 1)

    # 3b. the job loops after construction (optional: MUCHSALSA_GPU_STAGES)
    def wrap(loop_start: str, loop_end: str, gpu_call: str, occurrence: int = 0):
        nonlocal src
        i = -1
        for _ in range(occurrence + 1):
            i = src.index(loop_start, i + 1)
        j = src.index(loop_end, i) + len(loop_end)
        body = src[i:j]
        src = src[:i] + f"    if (gpuStages) {{\n      {gpu_call}\n    }} else {{\n" + body + "\n    }\n" + src[j:]

    wrap("    std::for_each(std::begin(edges), std::end(edges), [&](auto *const pEdge) {\n      wg.add(1);\n\n"
         "      auto job = Job(chainingJob,", "    wg.wait();\n",
         "muchsalsa::gpu::chainAndOverlaps(*pGpuCtx, graph, app.getWiggleRoom());")
    wrap("    std::for_each(std::begin(edges), std::end(edges), [&](auto const *const pEdge) {\n      wg.add(1);\n\n"
         "      auto job = Job(contractionEdgesJob,", "    wg.wait();\n",
         "muchsalsa::gpu::findContractionEdges(*pGpuCtx, graph, app.getWiggleRoom(), contractionEdges);")
    wrap("    std::for_each(std::begin(edges), std::end(edges), [&](auto *const pEdge) {\n      wg.add(1);\n"
         "      auto job = Job(deletableEdgesJob,", "    wg.wait();\n",
         "muchsalsa::gpu::findDeletableEdges(*pGpuCtx, graph, deletableEdges);")
    wrap("    std::for_each(std::begin(edges), std::end(edges), [&](auto *const pEdge) {\n      wg.add(1);\n"
         "      auto job = Job(computeBitweightJob,", "    wg.wait();\n",
         "muchsalsa::gpu::computeBitweights(*pGpuCtx, graph);")
    return src


def _inc(name: str) -> str:
    return open(os.path.join(HERE, "binary_patch", name)).read()


def patch_sequence_utils(src: str) -> str:
    """while a SequenceBatch is active on the calling thread, the two getters forward to it."""
    assert '#include "SequenceUtils.h"' in src
    src = src.replace('#include "SequenceUtils.h"', '#include "SequenceUtils.h"\n\n#include "ms_gpu_shim.h"', 1)
    for which, name in ((0, "getIlluminaSequence"), (1, "getNanoporeSequence")):
        m = re.search(r"std::string muchsalsa::%s\(([^)]*)\)\s*\{\n" % name, src)
        assert m, f"{name} has moved"
        args = [a.strip().split()[-1].lstrip("&*") for a in m.group(1).split(",")]  # parameter names, in order
        call = ("  if (auto *const pBatch = muchsalsa::gpu::SequenceBatch::active()) {\n"
                f"    return pBatch->get({which}U, {args[1]}, {args[2]}, {args[3]}, {args[4]});\n  }}\n")
        src = src[:m.end()] + call + src[m.end():]
    return src


def patch_main_sequences(src: str) -> str:
    """MUCHSALSA_GPU_SEQ=1 (with MUCHSALSA_GPU): the sequence files go to the resident store and every assemblePath call
    fetches what it needs as ONE batch (record pass -> msgpu_seq_fetch_batch -> replay pass; SequenceBatch)."""
    old = "    sequenceAccessor.buildIndex();\n"
    assert old in src
    src = src.replace(old, old + _inc("load_sequences_gpu.inc"), 1)
    decl = "void assemblePathsSub(gsl::not_null<Job const *> pJob);"
    assert decl in src
    src = src.replace(decl, decl + "\n\n#include <cstdio>\nstatic muchsalsa::gpu::Context *g_pGpuSeqCtx = nullptr; "
                                  "// set when the sequence store is on the GPU", 1)
    old = "    Id2OverlapMap id2OverlapMap;\n    muchsalsa::assemblePath(\n          &id2OverlapMap,\n"
    assert old in src
    return src.replace(old, _inc("assemble_sub_gpu.inc"), 1)


def main():
    build = os.environ.get("MSBIN_BUILD", "/tmp/msbin_build")
    os.makedirs(build, exist_ok=True)
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)  # the reference's objects + the shim object
    for f in ("Application.cpp", "Application.h", "TrackingAllocator.cpp", "TrackingAllocator.h"):
        open(os.path.join(build, f), "w").write(open(os.path.join(REF, "src", f)).read())
    open(os.path.join(build, "main.cpp"), "w").write(
        patch_main_sequences(patch_main(open(os.path.join(REF, "src", "main.cpp")).read())))
    open(os.path.join(build, "SequenceUtils.cpp"), "w").write(
        patch_sequence_utils(open(os.path.join(REF, "libms", "src", "SequenceUtils.cpp")).read()))
    cc = open(os.path.join(REF, "libms", "src", "kernel", "cc.cpp")).read()
    cc, n = re.subn(r"(std::deque<muchsalsa::graph::Vertex const \*)const>", r"\1>", cc)
    assert n == 1
    open(os.path.join(build, "cc.cpp"), "w").write(cc)

    cxx = os.environ.get("CXX", "g++")
    flags = ["-std=c++20", "-O2", "-DNDEBUG", "-pthread", "-include", "algorithm", f"-I{HERE}/gsl_standin",
             f"-I{REF}/include", f"-I{REF}/include/ms", f"-I{PKG}/host", f"-I{ROOT}/include", f"-I{build}"]
    objs = []
    units = [(os.path.join(build, f), f.replace(".cpp", ".o")) for f in ("main.cpp", "Application.cpp",
                                                                          "TrackingAllocator.cpp", "cc.cpp",
                                                                          "SequenceUtils.cpp")]
    units += [(os.path.join(REF, "libms", "src", u + ".cpp"), "bin_" + u.replace("/", "_") + ".o")
              for u in ("kernel/mst", "kernel/dg", "kernel/lg", "kernel/ap", "OutputWriter", "Kernel", "Util")]
    procs = []
    for src, o in units:
        obj = os.path.join(build, o)
        objs.append(obj)
        procs.append(subprocess.Popen([cxx, *flags, "-c", src, "-o", obj]))
    if any(p.wait() for p in procs):
        raise SystemExit("compilation failed")
    ref_objs = [os.path.join(OUT, "obj", f) for f in sorted(os.listdir(os.path.join(OUT, "obj")))
                if f.endswith(".o") and f not in ("ref_driver.o", "SequenceUtils.o")]  # (the patched copy replaces it)
    exe = os.path.join(OUT, "muchsalsa")
    subprocess.run([cxx, "-pthread", "-o", exe, *objs, *ref_objs, f"-L{PKG}", "-lmsgpu",
                    "-Wl,-rpath,$ORIGIN/../../muchsalsa_b200"], check=True)
    print(exe)


if __name__ == "__main__":
    main()
