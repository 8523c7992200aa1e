"""cuobjdump -sass summary of libgpb200.so for profiles/ (runs in the build container, no GPU): per kernel, the counts of
the Blackwell-specific mnemonics (UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor,
UTCBAR = tcgen05.commit, DMMA = FP64 tensor mma) and an excerpt of the MMA issue stream of gemm_i8_kernel<2>."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "inference_tools_b200", "libgpb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCIMMA.2CTA", "UTCIMMA", "LDTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "DMMA", "SYNCS", "ELECT", "DFMA", "MUFU"]
per, cur, excerpt = collections.OrderedDict(), None, []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "").replace("gpb::", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    ins = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if not ins:
        continue
    text = ins.group(1)
    per[cur]["_total"] += 1
    for mn in MN:
        if re.search(r"\b" + re.escape(mn) + r"\b" if "." not in mn else re.escape(mn), text):
            if mn == "UTCIMMA" and "2CTA" in text:
                continue
            per[cur][mn] += 1
    if cur.startswith("gemm_i8_kernel<2>") and ("UTCIMMA" in text or "UTCBAR" in text or "UTMALDG" in text) and len(excerpt) < 24:
        excerpt.append(text)
out = ["# SASS summary of libgpb200.so (round 2)", "",
       "`cuobjdump -sass inference_tools_b200/libgpb200.so`, built with `-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`.",
       "Mnemonics: UTCIMMA(.2CTA) = tcgen05.mma kind::i8 (cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA),",
       "UTCBAR = tcgen05.commit, DMMA = mma.sync m8n8k4 f64.  Kernels without any of them are omitted.", "",
       "| kernel | instr | " + " | ".join(MN) + " |", "|---|---:|" + "---:|" * len(MN)]
for k, c in per.items():
    if any(c[m] for m in MN[:7]):
        out.append(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[m]) for m in MN) + " |")
out += ["", "First tensor-core / TMA instructions of `gemm_i8_kernel<2>` in program order:", "", "```"] + excerpt + ["```", ""]
open(os.path.join(ROOT, "profiles", "sass_r2.md"), "w").write("\n".join(out))
print("\n".join(out[:40]))
