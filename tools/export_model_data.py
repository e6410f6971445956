"""Export the shipped model DATA (not code) from the reference checkout into the package.

Reads  /root/reference/stan_models/{arma,PRMwCD}/*.json|*.params|model_config.json
Writes smc-nuts_b200/smcnuts/data/<model>/...

`PRMwCD.json` in the reference is truncated after `"phi": `; it is
repaired here by appending `1.0}`.  Only data vectors travel; no reference source is copied.
Run once in the build container (the GPU box has no /root/reference).
"""
import json
import shutil
from pathlib import Path

REF = Path("/root/reference/stan_models")
OUT = Path(__file__).resolve().parents[1] / "smc-nuts_b200" / "smcnuts" / "data"


def main():
    for name in ("arma", "PRMwCD"):
        out = OUT / name
        out.mkdir(parents=True, exist_ok=True)
        raw = (REF / name / f"{name}.json").read_text()
        try:
            data = json.loads(raw)
        except json.JSONDecodeError:
            data = json.loads(raw + " 1.0}")
        data["phi"] = 1.0
        (out / f"{name}.json").write_text(json.dumps(data))
        shutil.copyfile(REF / name / f"{name}.params", out / f"{name}.params")
        shutil.copyfile(REF / name / "model_config.json", out / "model_config.json")
        print(name, {k: (len(v) if isinstance(v, list) else v) for k, v in data.items()})


if __name__ == "__main__":
    main()
