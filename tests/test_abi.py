"""The C-ABI library loads on a GPU-less host, exports every symbol include/parvar_b200.h declares, and refuses
to compute without a device (no CPU fallback)."""
import re
from pathlib import Path

import pytest

from conftest import ROOT, has_cuda, pkg


def _declared_symbols():
    text = (ROOT / "include" / "parvar_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    rt = pkg("runtime")
    lib = ctypes.CDLL(str(rt.LIB_PATH))
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/parvar_b200.h but not exported"
    assert sorted(rt.ABI) == syms, "runtime.ABI and the header disagree"


def test_registry_and_versions():
    rt = pkg("runtime")
    lib = rt.load_library()
    assert lib.pv_abi_version() == 1
    assert rt.model_names() == ["simple_chain", "simple_pk", "icg_body_flat"]
    P = pkg()
    assert set(P.MODELS) == {"simple_chain", "simple_pk", "icg_body_flat"}
    assert all(Path(p).exists() for p in P.MODELS.values())


@pytest.mark.skipif(has_cuda(), reason="checks the behaviour WITHOUT a device")
def test_no_cpu_fallback():
    rt = pkg("runtime")
    with pytest.raises(rt.ParvarB200Error, match="no CPU fallback"):
        rt.Problem("simple_pk")
    assert rt.fma_peak_tflops() < 0
    pf = pkg("experiments.petab_factory")
    with pytest.raises(rt.ParvarB200Error):
        pf.ODESampleSimulator(pkg().MODELS["simple_pk"])


def test_host_mirror_without_gpu():
    """Sampling and model resolution are host logic and work anywhere."""
    import numpy as np
    pf = pkg("experiments.petab_factory")
    P = pkg()
    assert pf.resolve_model(P.MODELS["icg_body_flat"]) == "icg_body_flat"
    s = pf.create_samples_parameters({"MALE": {"CL": pf.LognormParameters(n=20, sigma=1.0, mu=1.0),
                                               "k_abs": pf.LognormParameters(n=20, sigma=1.0, mu=0.5)}})
    np.testing.assert_allclose(s["MALE"]["k_abs"][:3], [0.81277741, 0.65610198, 0.97999947], rtol=1e-7)
    s2 = pf.create_samples_parameters({"MALE": {"k_abs": pf.LognormParameters(n=4, sigma=1.0, mu=0.5)}}, literal=False)
    assert abs(np.log(s2["MALE"]["k_abs"]).mean()) < 5      # intended parameterisation: just runs
    assert pf.lognorm_par_transform(1, 1) == pytest.approx((-0.3465735902799726, 0.8325546111576977))
    opt = pkg("optimization.petab_optimization")
    y = np.array([[[1.0, np.nan], [2.0, 3.0]], [[3.0, 1.0], [4.0, 5.0]]])
    st = opt.sufficient_statistics(y, 0)
    assert st.shape == (3, 2, 2) and st[0, 0, 1] == 1 and st[1, 0, 0] == 4 and st[2, 1, 1] == 34


def test_numa_helper_parses_cpulists_and_degrades_without_a_gpu():
    import importlib
    numa = importlib.import_module("parameter-variability_b200.numa")
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    info = numa.gpu_numa_info(0)
    assert set(info) == {"bus", "node", "cpus", "allowed"}
    import torch
    if not torch.cuda.is_available():
        assert numa.bind_to_gpu_node(0) is None     # nothing to bind to: the process affinity is left alone


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the header's structs (sampler options, MT19937 state, HMC / NUTS state, NUTS uniforms) have the
    C compiler's size and field offsets, and the NUTS slot lists follow the header's enums."""
    import ctypes
    import shutil
    import subprocess
    rt = pkg("runtime")
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    pairs = {"pv_sampler_options": rt.SamplerOptions, "pv_mt19937_state": rt.MT19937State, "pv_hmc_state": rt.HMCState,
             "pv_nuts_state": rt.NUTSState, "pv_nuts_uniforms": rt.NUTSUniforms}
    lines = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{ROOT / "include" / "parvar_b200.h"}"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines.append(f'  printf("enums %d %d %d %d\\n", PV_NUTS_NVEC, PV_NUTS_NSC, PV_NUTS_NIC, PV_NUTS_MAX_CHAINS);')
    lines.append(f'  printf("slots %d %d %d %d %d\\n", PV_NUTS_IM, PV_NUTS_GN, PV_NUTS_EPS, PV_NUTS_GROWING, PV_NUTS_DONE);')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run([cc, "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line, (cname, cls) in zip(out, pairs.items()):
        tok = line.split()
        assert tok[0] == cname
        assert int(tok[1]) == ctypes.sizeof(cls), cname
        assert [int(x) for x in tok[2:]] == [getattr(cls, f).offset for f, _ in cls._fields_], cname
    assert out[-2].split()[1:] == [str(len(rt.NUTS_VEC)), str(len(rt.NUTS_SC)), str(len(rt.NUTS_IC)), str(rt.NUTS_MAX_CHAINS)]
    assert out[-1].split()[1:] == [str(rt.NUTS_VEC.index("IM")), str(rt.NUTS_VEC.index("GN")), str(rt.NUTS_SC.index("EPS")),
                                   str(rt.NUTS_IC.index("GROWING")), str(rt.NUTS_IC.index("DONE"))]
