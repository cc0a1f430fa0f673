"""Warp-stall samples of an .ncu-rep aggregated per CUDA source line (needs -lineinfo + --import-source on).

    python profiles/by_line.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import subprocess
import sys


def main(path, top=60):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, H, per = None, None, collections.Counter()
    text, inst = {}, collections.Counter()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            H = {h: i for i, h in enumerate(r)}
        elif H and len(r) > 8 and r[0].strip().isdigit():  # a CUDA source line (SASS rows have an empty line number)
            key = (cur, int(r[0]))
            num = lambda x: int(x) if x.strip().isdigit() else 0
            per[key] += num(r[H["# Samples"]])
            inst[key] += num(r[H["Instructions Executed"]])
            text[key] = r[1].strip()[:90]
    tot = sum(per.values())
    files = collections.Counter()
    for (f, _), v in per.items():
        files[f] += v
    print(f"# {tot} samples")
    for f, v in files.most_common():
        print(f"{100 * v / tot:6.2f}%  {f}")
    print("# top lines")
    for (f, ln), v in per.most_common(top):
        print(f"{100 * v / tot:6.2f}%  {inst[(f, ln)]:>12d} inst  {f}:{ln}  {text[(f, ln)]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
