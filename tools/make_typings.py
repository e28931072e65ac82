"""Regenerates loco_hd_b200/primitive_typings/*.config.json from the reference's typing tables
(/root/reference/primitive_typings/*.config.json) in a canonical layout: one rule per line and the atom count
always explicit ([residue_regex, atom_regex, count]; the reference's two-element rules mean count = 1,
loco_hd/atom_converter_utils.py:79-81 upstream).  The tables are data that a drop-in must reproduce exactly:
type names, rule order and regexes are unchanged.  Run in the authoring container only."""
import json
import sys
from pathlib import Path

SRC = Path("/root/reference/primitive_typings")
DST = Path(__file__).resolve().parent.parent / "loco_hd_b200" / "primitive_typings"

for src in sorted(SRC.glob("*.config.json")):
    table = json.loads(src.read_text())
    lines = ["{"]
    items = list(table.items())
    for ti, (ptype, rules) in enumerate(items):
        lines.append(f"  {json.dumps(ptype)}: [")
        for ri, rule in enumerate(rules):
            full = [rule[0], rule[1], rule[2] if len(rule) == 3 else 1]
            lines.append("    " + json.dumps(full) + ("," if ri + 1 < len(rules) else ""))
        lines.append("  ]" + ("," if ti + 1 < len(items) else ""))
    lines.append("}")
    (DST / src.name).write_text("\n".join(lines) + "\n")
    print(src.name, {k: len(v) for k, v in table.items()}, file=sys.stderr)
