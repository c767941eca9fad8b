"""make_config_fixture -- TEST INFRASTRUCTURE ONLY.  Executes the reference's config files
(projects/configs/coocc_nusc/*.py are plain Python: assignments only) and stores their `model = dict(...)` entry as
JSON under tests/golden/configs/, so that the GPU box (which has no /root/reference) can build the detectors from the
*literal* configs.  Run in the build container only:

    python -m oracle.make_config_fixture
"""
import json
import os

REF = os.environ.get("COOCC_REFERENCE_ROOT", "/root/reference")
CFG_DIR = os.path.join(REF, "projects", "configs", "coocc_nusc")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "configs")


def load_model_dict(path):
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)      # noqa: S102 -- the reference's own config, assignments only
    return ns["model"]


def main():
    os.makedirs(OUT, exist_ok=True)
    for fn in sorted(os.listdir(CFG_DIR)):
        if not fn.endswith(".py"):
            continue
        model = load_model_dict(os.path.join(CFG_DIR, fn))
        out = os.path.join(OUT, fn[:-3] + ".model.json")
        with open(out, "w") as f:
            json.dump(dict(source="projects/configs/coocc_nusc/" + fn, model=model), f, indent=1, sort_keys=True)
        print(fn, "->", out, model["type"], sorted(k for k, v in model.items() if isinstance(v, dict) and "type" in v))


if __name__ == "__main__":
    main()
