"""`deep3d_aerial_b200.install()` -- the one-line binding of INTEGRATION.md §2.

  * against a stand-in `models` package written to a temp dir (runs everywhere): every hot-path name and
    forward is rebound, everything else is left alone;
  * against the live reference checkout when it is present (authoring container only): the reference's own
    classes are rebound, its networks still build, and a forward pass on CPU tensors now ends in this
    engine's "no CPU fallback" error -- i.e. the reference's call path really lands in the shim.
"""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/mvs/mvs_cas"

FAKE_MODULE = '''
def homo_warping_float(src_fea, src_proj, ref_proj, depth_values): return "upstream"
def homo_warping_double(src_fea, src_proj, ref_proj, depth_values): return "upstream"
def depth_regression(p, depth_values): return "upstream"
def get_depth_range_samples(*a, **k): return "upstream"
def get_cur_depth_range_samples(*a, **k): return "upstream"
def conv_block(x): return "untouched"
'''
FAKE_CAS = '''
from .module import *
class DepthNet:
    def forward(self, *a, **k): return "upstream"
class CostRegNet:
    def forward(self, x): return "untouched"
'''
FAKE_ADA = '''
from .module import *
class InferDepthNet:
    def forward(self, *a, **k): return "upstream"
class DepthNet:
    def forward(self, *a, **k): return "upstream"
'''


def _run(code, cwd=None, extra_path=()):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, *extra_path, os.environ.get("PYTHONPATH", "")]))
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env, cwd=cwd)


def test_install_rebinds_a_models_package(tmp_path):
    pkg = tmp_path / "fakeref" / "models"
    pkg.mkdir(parents=True)
    (pkg / "__init__.py").write_text("")
    (pkg / "module.py").write_text(FAKE_MODULE)
    (pkg / "cas_mvsnet.py").write_text(FAKE_CAS)
    (pkg / "adamvs.py").write_text(FAKE_ADA)
    r = _run('''
        import deep3d_aerial_b200 as d3d
        from deep3d_aerial_b200 import depthnets, module
        done = d3d.install()
        import models.module as m, models.cas_mvsnet as c, models.adamvs as a
        assert m.homo_warping_float is module.homo_warping_float and c.homo_warping_float is module.homo_warping_float
        assert a.depth_regression is module.depth_regression and c.get_depth_range_samples is module.get_depth_range_samples
        assert c.DepthNet.forward is depthnets.cas_depthnet_forward
        assert a.InferDepthNet.forward is depthnets.ada_infer_forward and a.DepthNet.forward is depthnets.ada_depthnet_forward
        assert m.conv_block(0) == "untouched" and c.CostRegNet().forward(0) == "untouched"
        assert "cas_mvsnet.DepthNet.forward" in done and "module.homo_warping_float" in done
        print("ok", len(done))
    ''', extra_path=[str(tmp_path / "fakeref")])
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_install_rebinds_the_live_reference():
    r = _run('''
        import torch
        import deep3d_aerial_b200 as d3d
        from deep3d_aerial_b200 import depthnets, module
        done = d3d.install()
        from models import module as m, cas_mvsnet, adamvs, msrednet, ucsnet
        assert cas_mvsnet.homo_warping_float is module.homo_warping_float
        assert cas_mvsnet.DepthNet.forward is depthnets.cas_depthnet_forward
        assert adamvs.InferDepthNet.forward is depthnets.ada_infer_forward
        assert msrednet.InferDepthNet.forward is depthnets.red_infer_forward
        assert ucsnet.compute_depth is depthnets.ucs_compute_depth
        net = cas_mvsnet.Infer_CascadeMVSNet(num_depth=48, ndepths=[8, 4, 2], depth_intervals_ratio=[4, 2, 1]).eval()
        keys = list(net.state_dict().keys())
        assert any(k.startswith("cost_regularization.0.") for k in keys) and any(k.startswith("feature.") for k in keys)
        imgs = torch.rand(1, 3, 3, 64, 96)
        eye = torch.eye(4).repeat(1, 3, 1, 1)
        proj = {"stage1": eye.clone(), "stage2": eye.clone(), "stage3": eye.clone()}
        try:
            with torch.no_grad():
                net(imgs, proj, torch.tensor([[5.0, 15.0]]))
        except RuntimeError as e:
            assert "no CPU fallback" in str(e), e
            print("ok", len(done))
        else:
            raise SystemExit("the reference's forward did not reach the engine")
    ''', extra_path=[REF])
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_predict_driver_builds_the_live_reference_networks(tmp_path):
    """`python -m deep3d_aerial_b200.predict --reference_root ...`: the reference's own network classes are
    built with predict.py's constructor arguments, their hot path is this engine's, their checkpoints' keys are
    untouched, and a `module.`-prefixed (nn.DataParallel) checkpoint loads."""
    r = _run('''
        import torch
        from deep3d_aerial_b200 import predict, depthnets
        for model, stage_key in (("adamvs", "DepthNet.0.reg_fuse."), ("casmvsnet", "cost_regularization.0."),
                                 ("msrednet", "cost_regularization.0.")):
            args = predict.build_parser().parse_args(["--data_folder", "x", "--output_folder", "y", "--model", model,
                                                      "--numdepth", "48", "--reference_root", %r])
            net = predict.build_model(args)
            keys = list(net.state_dict().keys())
            assert any(k.startswith("feature.") for k in keys) and any(k.startswith(stage_key) for k in keys), (model, keys[:5])
            if model == "adamvs":
                from models import adamvs
                assert adamvs.InferDepthNet.forward is depthnets.ada_infer_forward
                ckpt = {"model": {"module." + k: v for k, v in net.state_dict().items()}}
                torch.save(ckpt, "ckpt.pt")
                predict.load_checkpoint(net, "ckpt.pt")
        print("ok")
    ''' % REF, cwd=str(tmp_path))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
