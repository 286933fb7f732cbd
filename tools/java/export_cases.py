#!/usr/bin/env python
"""Exports the golden cases (tests/golden/*.npz) as Starfish input decks + case.txt for tools/java/ParityDump.java.
Doubles travel as the hex of their raw IEEE bits, so nothing is lost in text."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from starfish_b200 import synthetic as S  # noqa: E402

BC_NAMES = {"periodic": ["PERIODIC"] * 4, "open": ["OPEN"] * 4, "symmetry": ["SYMMETRY"] * 4,
            "beam": ["OPEN", "OPEN", "SYMMETRY", "OPEN"]}  # RIGHT, TOP, LEFT, BOTTOM (Mesh.Face.val())


def hexd(a):
    return " ".join("%x" % v for v in np.ascontiguousarray(a, np.float64).view(np.uint64).ravel())


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    for fn in sorted(os.listdir(gold)):
        if not fn.endswith(".npz"):
            continue
        name, z = fn[:-4], np.load(os.path.join(gold, fn))
        dom, ni, nj, n, steps = [int(v) for v in z["meta"]]
        d = os.path.join(gold, "java", name)
        os.makedirs(d, exist_ok=True)
        amu = float(z["mass"]) / S.AMU
        open(os.path.join(d, "starfish.xml"), "w").write(
            "<simulation>\n<note>parity case %s</note>\n<load>domain.xml</load>\n<load>materials.xml</load>\n"
            "<time><num_it>0</num_it><dt>%r</dt></time>\n<parity_dump case=\"case.txt\" out=\"out.txt\" />\n</simulation>\n" % (name, float(z["dt"])))
        open(os.path.join(d, "domain.xml"), "w").write(
            "<domain type=\"%s\">\n<mesh type=\"uniform\" name=\"mesh1\">\n<origin>0,0</origin>\n<spacing>1e-3,1e-3</spacing>\n"
            "<nodes>%d,%d</nodes>\n</mesh>\n</domain>\n" % (["xy", "rz", "zr"][dom], ni, nj))
        open(os.path.join(d, "materials.xml"), "w").write(
            "<materials>\n<material name=\"ion\" type=\"kinetic\">\n<molwt>%r</molwt>\n<charge>%r</charge>\n<spwt>1</spwt>\n</material>\n</materials>\n"
            % (amu, float(z["charge"]) / S.QE))
        with open(os.path.join(d, "case.txt"), "w") as f:
            f.write("steps %d\n" % steps)
            f.write("bc %s\n" % " ".join(BC_NAMES[str(z["bc"])]))
            f.write("efi %s\nefj %s\n" % (hexd(z["efi"]), hexd(z["efj"])))
            f.write("particles %d\n" % n)
            cols = np.stack([z["in_" + k] for k in ("x", "y", "z", "u", "v", "w", "mpw")], axis=1)
            for row in cols:
                f.write(hexd(row) + "\n")
        print("exported", name)


if __name__ == "__main__":
    main()
