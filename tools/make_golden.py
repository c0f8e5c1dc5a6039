#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ and the synthetic
SMPL template under jafpro_b200/data/ FROM THE REFERENCE TREE.

Run in the build container only (needs /root/reference, no GPU):

    python tools/make_golden.py

Everything written here is DATA produced by importing / executing the reference's
own Python modules on CPU (or by reading its golden images); no reference source
is copied.  The GPU box has no /root/reference, so tests read only the fixtures.

Fixtures (all small, float32 unless noted):
  teapot_faces.npz     faces [4928,3,3] after NR load_obj -> look_at -> perspective ->
                       vertices_to_faces with fill_back (tests/test_rasterize_silhouettes.py:16-35
                       path through renderer.py:74-95) + the Blender golden silhouette
                       tests/data/teapot_blender.png as packed bits.
  look_at.npz          the three known answers of tests/test_look_at.py:9-25 re-evaluated
                       with the reference look_at + SMPLRenderer-eye outputs on random vertices.
  render_faces.npz     cam, verts -> `faces` exactly as SMPLRenderer.render_fim_wim builds them
                       (src/nmr.py:263-276, rasteriser stubbed out).
  bc_transform.npz     SMPLRenderer.cal_bc_transform (src/nmr.py:617-659) on a random fim/wim.
  iuv_preprocessing.npz  TransferTexture (src/utils.py:369-394) and compute_angle (src/computer_angle.py:4-39) on synthetic IUV maps.
  vis_f2pts.npz        SMPLRenderer.get_vis_f2pts (src/nmr.py:507-546) on random fims (with / without background).
  convlstm.npz         src/convLSTM.py ConvLSTMCell.forward and a 3-step ConvLSTM.forward.
  softmax_fuse.npz     src/networks.py Downsampler_mask.forward K-reduction (:1259-1286), captured
                       with forward hooks at the first scale.
  mask_blend.npz       src/flow_net.py Propagation3DFlowNet.forward (:87-99).
  texture_warp.npz     test/conv_pro_test.py texture_warp_pytorch (:41-74), the IUV texture lookup.
  dataset_item.json    src/data.py Fusion_dataset_smpl_test.__getitem__ (:471-602) on a synthetic video written in the
                       reference's on-disk layout: shapes, dtypes and SHA-256 of every returned array.
  get_texture.npz      src/utils.py get_texture (:232-255) with this container's OpenCV (cv2.resize INTER_LINEAR on float64).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "jafpro_b200", "data")


def _import_reference():
    """Import reference modules on CPU: the three compiled CUDA extensions are absent,
    so empty stand-ins are registered for them (none is called), and `.cuda()` is a
    no-op because this container has no GPU."""
    import importlib.abc
    import importlib.machinery

    class _StubCuda(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, name, path, target=None):
            if name == "neural_renderer.cuda" or name.startswith("neural_renderer.cuda."):
                return importlib.machinery.ModuleSpec(name, self, is_package=(name == "neural_renderer.cuda"))
            return None

        def create_module(self, spec):
            return None

        def exec_module(self, module):
            pass

    sys.meta_path.insert(0, _StubCuda())
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, os.path.join(REF, "third_party", "neural_renderer"))
    sys.path.insert(0, REF)
    import neural_renderer as nr
    return nr


def teapot(nr):
    import cv2
    data = os.path.join(REF, "third_party", "neural_renderer", "tests", "data")
    vertices, faces = nr.load_obj(os.path.join(data, "teapot.obj"))
    assert faces.shape[0] == 2464 and vertices.shape[0] == 1292  # tests/test_load_obj.py:37-41
    vertices, faces = vertices[None], faces[None]
    # renderer.py:74-95 (render_silhouettes, camera_mode='look_at', fill_back=True, perspective)
    import math
    eye = [0, 0, -(1. / math.tan(math.radians(30)) + 1)]
    faces = torch.cat((faces, faces[:, :, list(reversed(range(faces.shape[-1])))]), dim=1)
    v = nr.look_at(vertices, eye)
    v = nr.perspective(v, angle=30)
    f3 = nr.vertices_to_faces(v, faces)[0].numpy().astype(np.float32)
    ref = cv2.imread(os.path.join(data, "teapot_blender.png"), cv2.IMREAD_UNCHANGED)
    ref = cv2.cvtColor(ref, cv2.COLOR_BGRA2RGBA) if ref.shape[-1] == 4 else ref
    sil = (ref.min(-1) != 255)  # test_rasterize_silhouettes.py:31-32
    depth_ref = cv2.imread(os.path.join(data, "test_depth.png"), cv2.IMREAD_UNCHANGED)
    np.savez_compressed(os.path.join(GOLD, "teapot_faces.npz"), faces=f3,
                        silhouette_bits=np.packbits(sil), shape=np.array(sil.shape),
                        depth_png=depth_ref.astype(np.uint8))
    print("teapot:", f3.shape, "covered", int(sil.sum()))


def look_at(nr):
    import math
    eyes = [[1, 0, 1], [0, 0, -10], [-1, 1, 0]]
    v1 = torch.from_numpy(np.array([1, 0, 0], np.float32))[None, None, :]
    outs = [nr.look_at(v1, np.array(e, np.float32)).numpy().squeeze() for e in eyes]
    answers = np.array([[-np.sqrt(2) / 2, 0, np.sqrt(2) / 2], [1, 0, 10],
                        [0, np.sqrt(2) / 2, 3. / 2. * np.sqrt(2)]])
    assert np.allclose(np.stack(outs), answers)  # tests/test_look_at.py:9-25
    g = torch.Generator().manual_seed(0)
    verts = torch.randn(2, 257, 3, generator=g)  # not 3: torch.cross(up, z) without dim= (look_at.py:49) picks dim 0 when B == 3
    eye = [0, 0, -(1. / np.tan(np.radians(30)) + 1)]  # src/nmr.py:177
    out = nr.look_at(verts.clone(), eye).numpy()
    np.savez_compressed(os.path.join(GOLD, "look_at.npz"), known_in=v1.numpy(), known_eyes=np.array(eyes, np.float32),
                        known_out=np.stack(outs), verts=verts.numpy(), smpl_eye_out=out)
    print("look_at ok; identity-rotation exact:", np.array_equal(out, (verts - torch.tensor(eye, dtype=torch.float32)).numpy()))


def render_faces(nr, tmpl_v, tmpl_f):
    import src.nmr as nmr
    captured = {}

    def fake_raster(faces, image_size, aa):
        captured["faces"] = faces.clone()
        return None, None

    nr.rasterize_face_index_map_and_weight_map = fake_raster
    rng = np.random.default_rng(1)
    B = 2
    verts = np.stack([tmpl_v + rng.normal(0, 2e-3, tmpl_v.shape) for _ in range(B)]).astype(np.float32)
    cam = np.stack([[0.83, 0.02, 0.27], [0.91, -0.04, 0.22]]).astype(np.float32)
    self = types.SimpleNamespace(faces=torch.tensor(tmpl_f.astype(np.int32)).int(),
                                 proj_func=nmr.orthographic_proj_withz_idrot, image_size=64,
                                 eye=[0, 0, -(1. / np.tan(np.radians(30)) + 1)])
    faces, _, _ = nmr.SMPLRenderer.render_fim_wim(self, torch.from_numpy(cam), torch.from_numpy(verts))
    sub = np.arange(0, tmpl_f.shape[0], 7)
    np.savez_compressed(os.path.join(GOLD, "render_faces.npz"), cam=cam, verts=verts,
                        face_subset=sub.astype(np.int32), faces_xyz_subset=faces.numpy()[:, sub])
    print("render_faces:", faces.shape)
    return nmr


def bc_transform(nmr):
    rng = np.random.default_rng(2)
    B, F, S = 2, 50, 16
    src = rng.normal(0, 0.5, (B, F, 3, 2)).astype(np.float32)
    fim = rng.integers(-1, F, (B, S, S)).astype(np.int32)
    w = rng.random((B, S, S, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    w[fim == -1] = 0
    self = types.SimpleNamespace(image_size=S)
    T = nmr.SMPLRenderer.cal_bc_transform(self, torch.from_numpy(src), torch.from_numpy(fim), torch.from_numpy(w))
    np.savez_compressed(os.path.join(GOLD, "bc_transform.npz"), src=src, fim=fim, wim=w, T=T.numpy())
    print("bc_transform:", T.shape)


def convlstm():
    from src.convLSTM import ConvLSTMCell, ConvLSTM
    torch.manual_seed(0)
    B, Cin, Ch, H, W = 2, 5, 6, 9, 7
    cell = ConvLSTMCell((H, W), Cin, Ch, (3, 3), True)
    x, h, c = torch.randn(B, Cin, H, W), torch.randn(B, Ch, H, W), torch.randn(B, Ch, H, W)
    with torch.no_grad():
        h2, c2 = cell(x, (h, c))
    K = 3
    lstm = ConvLSTM((H, W), Cin, Ch, (3, 3), 1, batch_first=True, bias=True, return_all_layers=False)
    xs = torch.randn(B, K, Cin, H, W)
    zeros = [(torch.zeros(B, Ch, H, W), torch.zeros(B, Ch, H, W))]  # init_hidden :58-63 without .cuda()
    with torch.no_grad():
        lo, last = lstm(xs, zeros)
    np.savez_compressed(
        os.path.join(GOLD, "convlstm.npz"), x=x.numpy(), h=h.numpy(), c=c.numpy(),
        weight=cell.conv.weight.detach().numpy(), bias=cell.conv.bias.detach().numpy(),
        h_out=h2.numpy(), c_out=c2.numpy(), seq_x=xs.numpy(),
        seq_weight=lstm.cell_list[0].conv.weight.detach().numpy(),
        seq_bias=lstm.cell_list[0].conv.bias.detach().numpy(), seq_out=lo.numpy(),
        seq_h=last[0][0].numpy(), seq_c=last[0][1].numpy())
    print("convlstm:", h2.shape, lo.shape)


def softmax_fuse():
    from src.networks import Downsampler_mask
    torch.manual_seed(0)
    enc = [12, 24, 24, 24, 24, 48, 48, 96, 96]
    net = Downsampler_mask(3, enc).eval()
    cap = {}
    net.mask1[0].register_forward_hook(lambda m, i, o: cap.update(feat=i[0].detach().clone(), logits=o.detach().clone()))
    xs = [torch.randn(2, 3, 32, 32) for _ in range(3)]
    with torch.no_grad():
        outs = net(xs)
    np.savez_compressed(os.path.join(GOLD, "softmax_fuse.npz"), feat=cap["feat"].numpy(),
                        logits=cap["logits"].numpy(), out=outs[0].numpy())
    print("softmax_fuse:", cap["feat"].shape, cap["logits"].shape, outs[0].shape)


def mask_blend():
    from src.flow_net import Propagation3DFlowNet
    torch.manual_seed(0)
    net = Propagation3DFlowNet(6, 8, 2, 1).eval()
    B, H, W = 2, 32, 32
    fake, tsf = torch.rand(B, 3, H, W) * 2 - 1, torch.rand(B, 3, H, W) * 2 - 1
    mask = (torch.rand(B, 1, H, W) > 0.4).float().repeat(1, 3, 1, 1)  # data.py:577-581: 3-ch mask
    with torch.no_grad():
        out = net({'fake_tgt': fake, 'tsf_image': tsf, 'tgt_IUV': None, 'use_IUV': False,
                   'use_mask': True, 'tgt_smpl_mask': mask})
    np.savez_compressed(os.path.join(GOLD, "mask_blend.npz"), fake=fake.numpy(), tsf=tsf.numpy(),
                        mask=mask.numpy(), weight=out['weight'].numpy(), pred=out['pred_target'].numpy())
    print("mask_blend:", out['pred_target'].shape)


def texture_warp():
    """test/conv_pro_test.py:41-74 texture_warp_pytorch, executed from the reference script file itself
    (the function is lifted out with ast because importing the script runs its CLI and needs a GPU)."""
    import ast
    src = open(os.path.join(REF, "test", "conv_pro_test.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "texture_warp_pytorch"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "conv_pro_test.py", "exec"), ns)
    rng = np.random.default_rng(0)
    H = W = 48
    Ht = Wt = 20
    tex = [torch.from_numpy(rng.normal(size=(3, Ht, Wt)).astype(np.float32)) for _ in range(24)]
    iuv = np.zeros((H, W, 3), np.uint8)
    iuv[..., 0] = rng.integers(0, 25, (H, W))
    iuv[..., 1] = rng.integers(0, 256, (H, W))
    iuv[..., 2] = rng.integers(0, 256, (H, W))
    iuv[0, :8, 1], iuv[0, :8, 2], iuv[1, :8, 1], iuv[1, :8, 2] = 0, 255, 255, 0  # texture corners
    out = ns["texture_warp_pytorch"](tex, iuv, "cpu")
    np.savez_compressed(os.path.join(GOLD, "texture_warp.npz"), tex=np.stack([t.numpy() for t in tex]), iuv=iuv,
                        out=out.numpy())
    print("texture_warp:", tuple(out.shape))


def vis_f2pts(nmr):
    """SMPLRenderer.get_vis_f2pts (src/nmr.py:507-546): faces absent from a fim get coordinates -2.  Item 0 has
    background pixels, item 1 has none (then `unique()[1:]` also drops the lowest visible face: a quirk we restate)."""
    rng = np.random.default_rng(6)
    B, F, S = 2, 60, 12
    f2pts = rng.normal(0, 0.5, (B, F, 3, 2)).astype(np.float32)
    fim = rng.integers(-1, F // 2, (B, S, S)).astype(np.int32)
    fim[1][fim[1] == -1] = 7
    out = nmr.SMPLRenderer.get_vis_f2pts(torch.from_numpy(f2pts), torch.from_numpy(fim))
    np.savez_compressed(os.path.join(GOLD, "vis_f2pts.npz"), f2pts=f2pts, fim=fim, out=out.numpy())
    print("vis_f2pts:", out.shape, "invisible faces:", int((out[..., 0, 0] == -2).sum()))


def iuv_preprocessing():
    """The per-frame IUV preprocessing of src/data.py (§8f rank 4): TransferTexture (src/utils.py:369-394, called three
    times per frame at src/data.py:102-113) and compute_angle (src/computer_angle.py:4-39, src/data.py:504).  Both are
    numpy-only; the modules' unrelated imports (matplotlib, tensorflow, cv2, moviepy) are absent here and are stubbed."""
    for name in ("matplotlib", "matplotlib.pyplot", "tensorflow", "cv2", "moviepy", "moviepy.editor"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__getattr__ = lambda attr: None
            m.__path__ = []
            sys.modules[name] = m
    from src.computer_angle import compute_angle
    from src.utils import TransferTexture
    sys.path.insert(0, ROOT)
    from oracle.inputs import iuv_preprocessing_inputs
    iuv, tex, im = iuv_preprocessing_inputs()
    n = iuv.shape[0]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        angles = np.array([compute_angle(iuv[i]) for i in range(n)], np.float64)
    out_bg = np.stack([TransferTexture(tex, iuv[i], im[i]) for i in range(n)])
    out_nobg = np.stack([TransferTexture(tex, iuv[i]) for i in range(n)])
    ones = np.stack([TransferTexture(np.ones((800, 1200, 3), np.uint8), iuv[i]) for i in range(n)])  # src/data.py:108
    # inputs are regenerated by iuv_preprocessing_inputs() in the tests; only the reference's outputs are stored
    np.savez_compressed(os.path.join(GOLD, "iuv_preprocessing.npz"), angles=angles, out_bg=out_bg, out_nobg=out_nobg,
                        ones=np.packbits(ones[..., 0] != 0))
    assert all(np.array_equal(ones[..., 0], ones[..., c]) for c in (1, 2)) and set(np.unique(ones)) <= {0, 1}
    print("iuv_preprocessing: angles", np.round(angles, 3))


def get_texture_fixture():
    """get_texture (src/utils.py:232-255), executed from the reference file itself with the real OpenCV of this container
    (the function is lifted out with ast: importing src.utils needs matplotlib / tensorflow / moviepy)."""
    import ast
    import cv2
    src = open(os.path.join(REF, "src", "utils.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_texture"][0]
    ns = {"np": np, "cv2": cv2}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "utils.py", "exec"), ns)
    sys.path.insert(0, ROOT)
    from oracle.inputs import iuv_preprocessing_inputs
    iuv, _, im = iuv_preprocessing_inputs()
    small = np.stack([np.stack(ns["get_texture"](im[i], iuv[i], tex_size=8, final_size=25)) for i in (0, 3)])
    full = np.stack(ns["get_texture"](im[1], iuv[1]))                      # defaults: 32 -> 200
    np.savez_compressed(os.path.join(GOLD, "get_texture.npz"), small=small, full_sub=full[:, ::7, ::7].copy(),
                        full_sum=full.sum(axis=(1, 2, 3)), cv2_version=np.array(cv2.__version__))
    print("get_texture:", small.shape, full.shape, "non-empty parts", int((full.sum(axis=(1, 2, 3)) > 0).sum()))


def dataset_item():
    """Fusion_dataset_smpl_test.__getitem__ (src/data.py:471-602) executed on a synthetic video written to disk in the
    reference's own layout (PNG files + pose_shape.pkl); the fixture keeps shapes, dtypes and SHA-256 digests of every
    returned array (the arrays themselves are ~100 MB)."""
    import hashlib
    import pickle
    import tempfile
    import cv2
    sys.path.insert(0, ROOT)
    from oracle.inputs import synthetic_video
    for name in ("matplotlib", "matplotlib.pyplot", "tensorflow", "moviepy", "moviepy.editor"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__getattr__ = lambda attr: None
            m.__path__ = []
            sys.modules[name] = m
    if not hasattr(np, "int"):
        np.int = int  # the reference targets numpy < 1.24 (src/data.py:511)
    from src.data import Fusion_dataset_smpl_test
    v = synthetic_video()
    T = v["img"].shape[0]
    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        vid = os.path.join(tmp, "data", "test", "Synth_video_0_1")
        msk = os.path.join(tmp, "mask", "test", "Synth_video_0_1")
        smp = os.path.join(tmp, "smpl", "test", "Synth_video_0_1")
        for d_ in (vid, msk, smp, os.path.join(tmp, "log_result")):
            os.makedirs(d_)
        for t in range(T):
            cv2.imwrite(os.path.join(vid, f"frame_{t}.png"), v["img"][t])
            cv2.imwrite(os.path.join(vid, f"frame_{t}_IUV.png"), v["iuv"][t])
            cv2.imwrite(os.path.join(vid, f"frame_{t}_text.png"), v["text"][t])
            cv2.imwrite(os.path.join(vid, f"frame_{t}_mask.png"), v["text_mask"][t])
            cv2.imwrite(os.path.join(msk, f"frame_{t}_mask.png"), v["real_mask"][t])
        with open(os.path.join(smp, "pose_shape.pkl"), "wb") as fh:
            pickle.dump({k: v[k] for k in ("cams", "pose", "shape", "vertices")}, fh)
        for n_in in (4, 3, 1):
            ds = object.__new__(Fusion_dataset_smpl_test)   # __init__ only parses options: set what __getitem__ reads
            ds.vid_list = [vid]
            ds.smpl_dir, ds.mask_dir = os.path.join(tmp, "smpl", "test"), os.path.join(tmp, "mask", "test")
            ds.num_inputs, ds.output_mask = n_in, True
            ds.log_file_dir = os.path.join(tmp, "log_result", "chosen_frame_train.txt")
            src_data, tgt_data, data_255, smpl_data, vid_name, names, pro_frames = ds[0]
            flat = {"src_%d" % i: a for i, a in enumerate(src_data)}
            flat.update({"tgt_%d" % i: a for i, a in enumerate(tgt_data)})
            flat.update({"u255_%d" % i: a for i, a in enumerate(data_255)})
            flat.update({"smpl_%d" % i: a for i, a in enumerate(smpl_data)})
            for k, a in flat.items():
                a = np.ascontiguousarray(a)
                rec[f"n{n_in}/{k}"] = [str(a.dtype), list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()]
            rec[f"n{n_in}/pro_frames"] = [int(x) for x in pro_frames]
            rec[f"n{n_in}/vid_name"] = vid_name
            rec[f"n{n_in}/img_names"] = list(names)
    import json
    with open(os.path.join(GOLD, "dataset_item.json"), "w") as fh:
        json.dump(rec, fh, indent=1, sort_keys=True)
    print("dataset_item:", {k: v_ for k, v_ in rec.items() if k.endswith("pro_frames")})


def smpl_template():
    """mapper.txt `v` lines (6890 T-pose vertices) + smpl_faces.npy -> jafpro_b200/data/."""
    vs = []
    with open(os.path.join(REF, "mapper.txt")) as f:
        for line in f:
            if line.startswith("v "):
                vs.append([float(t) for t in line.split()[1:4]])
    v = np.array(vs, np.float32)
    faces = np.load(os.path.join(REF, "smpl_faces.npy"))
    assert v.shape == (6890, 3) and faces.shape == (13776, 3) and faces.max() == 6889
    os.makedirs(DATA, exist_ok=True)
    np.savez_compressed(os.path.join(DATA, "smpl_template.npz"), verts=v, faces=faces.astype(np.uint16))
    print("smpl_template:", v.shape, faces.shape, "bbox", v.min(0), v.max(0))
    return v, faces.astype(np.int32)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    nr = _import_reference()
    v, f = smpl_template()
    teapot(nr)
    look_at(nr)
    nmr = render_faces(nr, v, f)
    bc_transform(nmr)
    vis_f2pts(nmr)
    convlstm()
    softmax_fuse()
    mask_blend()
    texture_warp()
    get_texture_fixture()   # before iuv_preprocessing(): that one registers stand-ins for modules it does not need
    dataset_item()
    iuv_preprocessing()
