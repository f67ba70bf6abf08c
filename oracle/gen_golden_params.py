"""Golden vectors for SURVEY.md section 8 row a17 (utils.py:151-185 get_file_path, :2662-2741 get_specific_sae_params):
runs the REAL reference functions (extracted from /root/reference/utils.py, see ref_import.py) for every GoogLeNet
layer name and freezes their outputs in tests/golden/sae_params_table.json.  TEST INFRASTRUCTURE ONLY; needs
/root/reference, so it runs in the build container and its output is committed."""
import ast
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    with open(os.path.join(ref_import.REFERENCE_ROOT, "utils.py")) as fh:
        tree = ast.parse(fh.read())
    names = ("get_file_path", "get_specific_sae_params")
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"os": os}
    exec(compile(ast.Module(body=body, type_ignores=[]), "utils.py", "exec"), ns)
    model_params_temp = {"model_name": "inceptionv1", "epochs": "0", "lr": "0.001", "bs": "512", "opt": "sgd"}
    out = {"model_params_temp": [[k, v] for k, v in model_params_temp.items()], "layers": {}, "paths": []}   # ordered
    layers = [p + s for p in ("mixed", "inception") for s in ("3a", "3b", "4a", "4b", "4c", "4d", "4e", "5a", "5b")]
    for layer in layers:
        for sae, opt in (("sae_mlp", "constrained_adam"), ("gated_sae", "adam")):
            r = ns["get_specific_sae_params"](layer, sae, dict(model_params_temp), opt)
            out["layers"][f"{layer}|{sae}|{opt}"] = list(r)
    cases = [(None, "mixed3a", None, "x.pth", None), ("f", "mixed3a", "abc", ".pth", None),
             ("f", "mixed4c", {"a": 1, "b": None}, "model_weights.pth", None),
             ("f", "mixed4c", {"a": 1}, "w.pth", {"c": "z", "d": 2}), ("f", None, None, None, None)]
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        for folder, layer, params, fname, params2 in cases:
            fp = ns["get_file_path"](os.path.join(tmp, folder) if folder else None, layer, params, fname, params2)
            out["paths"].append({"folder": folder, "layer": layer, "params": params, "file_name": fname, "params2": params2,
                                 "result": os.path.relpath(fp, tmp) if folder else fp})
    path = os.path.join(ROOT, "tests", "golden", "sae_params_table.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(out["layers"]), "table rows")


if __name__ == "__main__":
    main()
