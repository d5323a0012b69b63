"""Extracts the only deterministic known-answer vector the reference holds for the hot path:
notebooks/modeling.ipynb cell 13 (source lines 16967-16979) plots |simple_flattopHG_field(x,0,0,p)|^2 and
|simple_flattopLG_field(x,0,0,p)|^2 + 0.001 for x = -2:0.01:2, p = [1, 1/sqrt(s), 5, tag, i, i], i in 0:2:8,
s = max(i,1); the 10 traces are embedded in the notebook output as plotly JSON.

Run in the build container (needs /root/reference); writes tests/golden/beam_profiles.json (committed).
"""
import json
import os

nb = json.load(open("/root/reference/notebooks/modeling.ipynb"))
cell = nb["cells"][13]
fig = [o for o in cell["outputs"] if "data" in o and "application/vnd.plotly.v1+json" in o["data"]][0]
data = fig["data"]["application/vnd.plotly.v1+json"]["data"]
out = {"source": "reference notebooks/modeling.ipynb cell 13", "x": data[0]["x"], "traces": {}}
for tr in data:
    assert tr["x"] == out["x"]
    out["traces"][tr["name"]] = tr["y"]
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "beam_profiles.json")
json.dump(out, open(dst, "w"))
print(sorted(out["traces"]), len(out["x"]))
