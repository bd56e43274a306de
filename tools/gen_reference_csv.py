#!/usr/bin/env python3
"""Run the reference's data generator reproducibly (authoring container only).

The reference's data-generation/generate_commands.py is unseeded and reads the wall clock
(generate_commands.py:627-635 `datetime.utcnow()`, `random` never seeded), and its two CSV
fixtures are git-LFS pointer files, so the data has to be regenerated.  This wrapper seeds
`random` and freezes `datetime.utcnow()` and then runs the UNMODIFIED script from
/root/reference with runpy -- nothing is copied.  It is how tests/golden/commands_2k.csv
was made:

    python tools/gen_reference_csv.py 2000 tests/golden/commands_2k.csv

/root/reference does not exist on the GPU box; there, CSVs come from our own generator
(`qpe_datagen`, csrc/datagen.cpp), which restates the distributions of SURVEY.md App. C.
"""
import datetime as _dt
import random
import runpy
import sys

REF_SCRIPT = "/root/reference/data-generation/generate_commands.py"
SEED = 12345
FROZEN_NOW = _dt.datetime(2026, 10, 18, 0, 0, 0)


class _FrozenDateTime(_dt.datetime):
    @classmethod
    def utcnow(cls):
        return cls(FROZEN_NOW.year, FROZEN_NOW.month, FROZEN_NOW.day)


def main():
    if len(sys.argv) < 3:
        print(f"usage: {sys.argv[0]} NUM_ROWS OUTPUT_CSV [SEED]", file=sys.stderr)
        return 2
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else SEED
    random.seed(seed)
    _dt.datetime = _FrozenDateTime  # the script does `from datetime import datetime`
    sys.argv = ["generate_commands.py", sys.argv[1], sys.argv[2]]
    runpy.run_path(REF_SCRIPT, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
