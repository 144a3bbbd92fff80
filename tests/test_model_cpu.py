"""CPU tests of the host-side model layer (no device calls): YAML contract and grid builders."""
import numpy as np
import pytest
import yaml

from common import ks_yaml_dict
from oracle import oracle as O


def test_build_model_from_yaml(tmp_path):
    from hankb200 import model as M
    p = tmp_path / "ks.yaml"
    p.write_text(yaml.safe_dump(ks_yaml_dict(), allow_unicode=True), encoding="utf-8")
    mod = M.build_model_from_yaml(str(p))
    assert mod.compspec.T == 150 and mod.compspec.eps == 1e-6 and mod.compspec.n_v == 6
    assert mod.var_names == ("Y", "KS", "r", "w", "KD", "Z")
    w, pr = mod.heterogeneity["wealth"], mod.heterogeneity["productivity"]
    assert w.n == 200 and pr.n == 7 and w.policy_var == "KD"
    assert np.allclose(w.grid, O.double_exponential(200, 0.0, 200.0), rtol=1e-14, atol=1e-16)
    zo, Pio, _ = O.rouwenhorst(7, 0.966, 0.283)
    assert np.allclose(pr.grid, zo, rtol=1e-15) and np.allclose(pr.transition, Pio, rtol=1e-15, atol=1e-18)
    assert mod.ss_initial["fixed"]["Z"] == 1.0 and mod.ss_ending["fixed"]["Z"] == 2.0
    mod2 = M.build_model_from_yaml(str(p), {"T": 300, "wealth.n": 500})
    assert mod2.compspec.T == 300 and mod2.heterogeneity["wealth"].n == 500


def test_unsupported_models_are_rejected(tmp_path):
    from hankb200 import model as M
    d = ks_yaml_dict(); d["equations"][0] = "Y = Z * KS(-1)^α + 1"
    p = tmp_path / "bad.yaml"; p.write_text(yaml.safe_dump(d, allow_unicode=True), encoding="utf-8")
    with pytest.raises(NotImplementedError):
        M.build_model_from_yaml(str(p))
    d = ks_yaml_dict(); d["variables"]["heterogeneous"][1]["function"] = "OtherValueFunction"
    p.write_text(yaml.safe_dump(d, allow_unicode=True), encoding="utf-8")
    with pytest.raises(NotImplementedError):
        M.build_model_from_yaml(str(p))
