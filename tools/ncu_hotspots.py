"""Per-source-line warp-stall map of a kernel from an .ncu-rep captured with `ncu --set full --import-source on` (needs the ncu CLI).

    python tools/ncu_hotspots.py gpurun_out/prof_gram_tc_v6_131072.ncu-rep [top=45] > profiles/<name>_hotspots.txt

Reads `ncu -i REP --page source --csv --print-source cuda,sass`, keeps the CUDA-line rows (an inlined helper is listed both at its
own line and at its call site, so percentages can sum to more than 100), and prints the lines with the most stall samples together
with their two dominant stall reasons."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur, hdr, kernel, out = None, None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) == 2 and r[0] == "Function Name":
            kernel = r[1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0]:
            try:
                out.append((int(r[hdr.index("# Samples")]), cur, int(r[0]), r[1].strip(), r))
            except ValueError:
                pass
    tot = sum(o[0] for o in out) or 1
    stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    print(f"# {kernel}\n# {rep}: {tot} samples over {len(out)} source lines; share, file:line, source, two dominant stall reasons")
    for s, f, ln, src, r in sorted(out, key=lambda o: -o[0])[:top]:
        st = sorted(((int(r[i]), h) for i, h in stalls), reverse=True)[:2]
        print(f"{100 * s / tot:5.1f}%  {f}:{ln:<4d} {src[:88]:88s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
    # shared-memory wavefronts per source line (the data pipe the tensor-core kernels are bound by): total, ideal, excessive
    cols = {h: i for i, h in enumerate(hdr)}
    wf = next((h for h in hdr if h.startswith("L1 Wavefronts Shared") and "Excessive" not in h and "Ideal" not in h), None)
    ex = next((h for h in hdr if h.startswith("L1 Wavefronts Shared Excessive")), None)
    if wf and ex:
        def num(r, h):
            try:
                return int(float(r[cols[h]].replace(",", "")))
            except ValueError:
                return 0
        rows_wf = sorted(((num(r, wf), num(r, ex), f, ln, src) for _, f, ln, src, r in out), reverse=True)
        twf = sum(x[0] for x in rows_wf) or 1
        print(f"# shared-memory wavefronts by source line ({twf} in total, {sum(x[1] for x in rows_wf)} excessive)")
        for w_, e_, f, ln, src in rows_wf[:20]:
            print(f"{100 * w_ / twf:5.1f}%  {f}:{ln:<4d} wavefronts={w_} excessive={e_}  {src[:80]}")


if __name__ == "__main__":
    main()
