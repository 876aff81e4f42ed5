"""Per-kernel SASS mnemonic table of the built objects (run here, no GPU needed): proves which kernels issue tcgen05 / TMA
instructions and how the MMAs are issued.  python tools/sass_table.py > profiles/<file>

Columns: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), of which `in R2UR.BROADCAST loop` = preceded within 8 instructions by the
ELECT + R2UR.BROADCAST sequence ptxas emits when the issuing lane is picked by `if (lane == 0)` instead of elect.sync in
uniform control flow; UTMALDG / UTMASTG / UTMAPF (TMA load / store / L2 prefetch), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
legacy HMMA (must be 0)."""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    print("# cuobjdump -sass of build/*.o (the objects libfacfake.so is linked from), one line per kernel that issues tensor-core or TMA instructions")
    print(f"# {'unit':12s} {'kernel':78s} UTCHMMA (.2CTA) in-R2UR.BROADCAST-loop | UTMALDG UTMASTG UTMAPF | LDTM UTCBAR | HMMA")
    for obj in sorted(glob.glob(os.path.join(ROOT, "build", "ff_*.o"))):
        unit = os.path.basename(obj)[:-2]
        if unit.endswith("_trace"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        for block in sass.split("Function : ")[1:]:
            mangled = block.split("\n", 1)[0].strip()
            name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            ins = [m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", block, re.M)]
            n = {k: sum(1 for i in ins if i.startswith(k)) for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCBAR", "HMMA")}
            if not (n["UTCHMMA"] or n["UTMALDG"] or n["UTMASTG"] or n["HMMA"]):
                continue
            two = sum(1 for i in ins if i.startswith("UTCHMMA") and "2CTA" in i)
            loop = sum(1 for k, i in enumerate(ins) if i.startswith("UTCHMMA") and any(j.startswith("R2UR.BROADCAST") for j in ins[max(0, k - 8):k]))
            name = re.sub(r"^void ", "", name).replace("ff::", "")
            print(f"  {unit:12s} {name[:78]:78s} {n['UTCHMMA']:3d} ({two:3d})  {loop:3d} | {n['UTMALDG']:3d} {n['UTMASTG']:3d} {n['UTMAPF']:3d} | {n['LDTM']:3d} {n['UTCBAR']:3d} | {n['HMMA']:3d}")


if __name__ == "__main__":
    main()
