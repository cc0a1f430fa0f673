"""Executed warp-instructions and stall samples of an .ncu-rep aggregated per source FILE and per kernels.cuh
line range (region), FP64 vs other opcodes.   python profiles/by_region.py gpurun_out/prof.ncu-rep"""
import collections, csv, re, subprocess, sys

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, H, curline = None, None, None
    inst, samp, fp64 = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            H = {h: i for i, h in enumerate(r)}
        elif H and len(r) > 8:
            num = lambda x: int(x) if x.strip().isdigit() else 0
            if r[0].strip().isdigit():
                curline = (cur, int(r[0]))
            else:  # SASS row under the current source line
                ie, sm = num(r[H["Instructions Executed"]]), num(r[H["# Samples"]])
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
                op = m.group(2).split(".")[0] if m else "?"
                inst[curline] += ie; samp[curline] += sm
                if op in ("DFMA", "DADD", "DMUL", "DSETP", "MUFU"): fp64[curline] += ie
    ti, ts = sum(inst.values()), sum(samp.values())
    files = collections.defaultdict(lambda: [0, 0, 0])
    for k in inst:
        f = files[k[0]]; f[0] += inst[k]; f[1] += samp[k]; f[2] += fp64[k]
    print(f"# {ti} warp-inst, {ts} samples")
    for f, (i, s, d) in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"{f:28s} inst {100*i/ti:6.2f}%  samples {100*s/ts:6.2f}%  fp64 share of its inst {100*d/max(i,1):5.1f}%")
    print("# top lines by executed instructions")
    for k, v in inst.most_common(40):
        print(f"{100*v/ti:6.2f}% inst {100*samp[k]/ts:6.2f}% samp fp64 {100*fp64[k]/max(v,1):5.1f}%  {k[0]}:{k[1]}")

main(sys.argv[1])
