"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` and by the
``not gpu`` tests that pin the oracle against the reference when
``/root/reference`` exists (it does not exist on the GPU box, where the
committed fixtures in ``tests/golden/`` stand in).

``diffusers`` is not installed here, and the reference's model / scheduler files
import four things from it (model.py:7-8, scheduling_unipc_multistep_clean.py:22-24).
A minimal stand-in for exactly those names is registered in ``sys.modules``;
nothing of the reference is copied or edited.  ``flash_attention`` asserts CUDA
(attention.py:54), so on the CPU it is replaced by an equivalent SDPA call.
"""
from __future__ import annotations

import functools
import importlib
import inspect
import os
import sys
import types
from enum import Enum

import torch

REF_ROOT = "/root/reference"
REF_WAN = os.path.join(REF_ROOT, "wan_for_worldforge")


def available() -> bool:
    return os.path.isdir(REF_WAN)


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _register_to_config(init):
    @functools.wraps(init)
    def wrapped(self, *a, **kw):
        sig = inspect.signature(init)
        bound = sig.bind(self, *a, **kw)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        prev = getattr(self, "_shim_cfg", None)
        self._shim_cfg = _Cfg(cfg if prev is None else {**prev, **cfg})
        init(self, *a, **kw)
    return wrapped


class _ConfigMixin:
    @property
    def config(self):
        return self._shim_cfg

    def register_to_config(self, **kw):
        self._shim_cfg.update(kw)

    @classmethod
    def from_config(cls, config, **kw):
        return cls(**{**dict(config), **kw})


class _ModelMixin(torch.nn.Module):
    pass


class _SchedulerMixin:
    pass


class _BaseOutput(dict):
    pass


class _Karras(Enum):
    UniPCMultistepScheduler = 1


def install_diffusers_shim() -> None:
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_wf_shim", False):
        return  # a real diffusers is present: use it
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    d = mod("diffusers"); d._wf_shim = True
    cu = mod("diffusers.configuration_utils")
    cu.ConfigMixin = _ConfigMixin
    cu.register_to_config = _register_to_config
    mm = mod("diffusers.models"); mu = mod("diffusers.models.modeling_utils")
    mu.ModelMixin = _ModelMixin
    ut = mod("diffusers.utils")
    ut.deprecate = lambda *a, **k: None
    ut.is_scipy_available = lambda: True
    ut.BaseOutput = _BaseOutput
    import logging as _logging
    ut.logging = types.SimpleNamespace(get_logger=_logging.getLogger)
    sc = mod("diffusers.schedulers"); su = mod("diffusers.schedulers.scheduling_utils")
    su.KarrasDiffusionSchedulers = _Karras
    su.SchedulerMixin = _SchedulerMixin
    su.SchedulerOutput = _BaseOutput
    d.configuration_utils, d.models, d.utils, d.schedulers = cu, mm, ut, sc
    mm.modeling_utils = mu
    sc.scheduling_utils = su


def _sdpa_flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None,
                          q_scale=None, causal=False, window_size=(-1, -1), deterministic=False,
                          dtype=torch.bfloat16, version=None):
    """CPU stand-in with flash_attention's dtype contract (attention.py:54-130):
    inputs that are not half are cast to ``dtype``; the result is cast back to q's dtype."""
    assert q_lens is None and not causal
    if k_lens is not None:   # self-attention passes the (full) sequence length
        assert all(int(n) == k.size(1) for n in k_lens)
    out_dtype = q.dtype
    half = (torch.float16, torch.bfloat16)
    cast = (lambda x: x if x.dtype in half else x.to(dtype)) if _FLASH_CAST[0] else (lambda x: x)
    q, k, v = cast(q), cast(k), cast(v)
    q = q.to(v.dtype); k = k.to(v.dtype)
    o = torch.nn.functional.scaled_dot_product_attention(
        q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float(),
        scale=softmax_scale).transpose(1, 2)
    if _FLASH_CAST[0]:
        o = o.to(v.dtype)
    return o.contiguous().type(out_dtype)


_FLASH_CAST = [False]   # False: keep fp32 end to end (fp32 pinning); True: flash-attn's bf16 casts


def _add_path():
    if REF_WAN not in sys.path:
        sys.path.insert(0, REF_WAN)


def load_wan_model_module(cpu_autocast: bool = False):
    """Returns wan.modules.model with flash_attention replaced for the CPU.

    cpu_autocast=True additionally maps the ``torch.cuda.amp.autocast`` the
    module uses for its fp32 islands (model.py:31,42,297,305,312,344,546) onto
    CPU autocast, so that running the module under
    ``torch.autocast('cpu', dtype=torch.bfloat16)`` reproduces the dtype flow it
    has on a GPU under ``torch.autocast('cuda', dtype=torch.bfloat16)``.
    """
    assert available()
    install_diffusers_shim()
    _add_path()
    if cpu_autocast:
        import torch.cuda.amp as camp
        if not getattr(camp, "_wf_patched", False):
            camp._orig_autocast = camp.autocast
            camp.autocast = lambda *a, **kw: torch.autocast("cpu", *a, **kw)
            camp._wf_patched = True
    for name in [m for m in sys.modules if m == "wan" or m.startswith("wan.")]:
        del sys.modules[name]
    # import the sub-modules directly; wan/__init__ pulls in pipelines we do not need
    pkg = types.ModuleType("wan"); pkg.__path__ = [os.path.join(REF_WAN, "wan")]
    sys.modules["wan"] = pkg
    mods = types.ModuleType("wan.modules"); mods.__path__ = [os.path.join(REF_WAN, "wan", "modules")]
    sys.modules["wan.modules"] = mods
    att = importlib.import_module("wan.modules.attention")
    att.flash_attention = _sdpa_flash_attention
    model = importlib.import_module("wan.modules.model")
    model.flash_attention = _sdpa_flash_attention
    _FLASH_CAST[0] = cpu_autocast
    return model


def load_wan_encoder_modules():
    """(wan.modules.t5, wan.modules.clip) of the reference - the UMT5 text encoder and the CLIP ViT image encoder - with the
    tokenizer's ``ftfy`` dependency stubbed (not installed; tokenisation is not on the path) and flash_attention replaced for
    the CPU."""
    load_wan_model_module()
    if "ftfy" not in sys.modules:
        f = types.ModuleType("ftfy"); f.fix_text = lambda t: t
        sys.modules["ftfy"] = f
    cur = torch.cuda.current_device                 # t5.py:478 evaluates it in a default argument at import time
    torch.cuda.current_device = lambda: 0
    try:
        t5 = importlib.import_module("wan.modules.t5")
        clip = importlib.import_module("wan.modules.clip")
    finally:
        torch.cuda.current_device = cur
    clip.flash_attention = _sdpa_flash_attention
    return t5, clip


def load_wan_vae_module():
    assert available()
    _add_path()
    if "wan" not in sys.modules:
        pkg = types.ModuleType("wan"); pkg.__path__ = [os.path.join(REF_WAN, "wan")]
        sys.modules["wan"] = pkg
        mods = types.ModuleType("wan.modules"); mods.__path__ = [os.path.join(REF_WAN, "wan", "modules")]
        sys.modules["wan.modules"] = mods
    return importlib.import_module("wan.modules.vae")


def load_scheduler_module():
    assert available()
    install_diffusers_shim()
    _add_path()
    if "utils" in sys.modules and not hasattr(sys.modules["utils"], "__path__"):
        del sys.modules["utils"]
    spec = importlib.util.spec_from_file_location(
        "wf_ref_scheduling_unipc", os.path.join(REF_WAN, "utils", "scheduling_unipc_multistep_clean.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


REF_LONGCAT = os.path.join(REF_ROOT, "longcat_for_worldforge")


def _fake_xformers():
    """xformers is not installed; the LongCat attention modules call it when ``enable_xformers`` is set.  A stand-in
    with the two entry points they use (attention.py:100, :243-246), computed with SDPA in fp32."""
    if "xformers" in sys.modules:
        return
    xf = types.ModuleType("xformers"); ops = types.ModuleType("xformers.ops")
    fmha = types.ModuleType("xformers.ops.fmha"); ab = types.ModuleType("xformers.ops.fmha.attn_bias")

    class BlockDiagonalMask:
        def __init__(self, q_lens, k_lens):
            self.q_lens, self.k_lens = q_lens, k_lens

        @classmethod
        def from_seqlens(cls, q_lens, k_lens):
            return cls(list(q_lens), list(k_lens))

    def memory_efficient_attention(q, k, v, attn_bias=None, op=None):
        # q [B, M, H, K]; with a block-diagonal mask B == 1 and the sequences are concatenated along M
        def sdpa(a, b, c):
            o = torch.nn.functional.scaled_dot_product_attention(a.transpose(1, 2).float(), b.transpose(1, 2).float(),
                                                                 c.transpose(1, 2).float())
            return o.transpose(1, 2).to(a.dtype)
        if attn_bias is None:
            return sdpa(q, k, v)
        outs, qo, ko = [], 0, 0
        for ql, kl in zip(attn_bias.q_lens, attn_bias.k_lens):
            outs.append(sdpa(q[:, qo:qo + ql], k[:, ko:ko + kl], v[:, ko:ko + kl]))
            qo += ql; ko += kl
        return torch.cat(outs, dim=1)

    ab.BlockDiagonalMask = BlockDiagonalMask
    fmha.attn_bias = ab
    ops.fmha = fmha
    ops.memory_efficient_attention = memory_efficient_attention
    xf.ops = ops
    sys.modules.update({"xformers": xf, "xformers.ops": ops, "xformers.ops.fmha": fmha, "xformers.ops.fmha.attn_bias": ab})


def load_longcat_bsa_module():
    """longcat_video.block_sparse_attention.bsa_interface, unmodified.  Its Triton kernels run on the CPU through Triton's
    interpreter (TRITON_INTERPRET=1, which must be set before triton is first imported) and its @torch.compile helpers
    eagerly (TORCHDYNAMO_DISABLE=1).  The interpreter has no bf16 (numpy), so fixtures from it are fp32 / fp16."""
    assert os.path.isdir(REF_LONGCAT)
    assert "triton" not in sys.modules or os.environ.get("TRITON_INTERPRET") == "1", "set TRITON_INTERPRET=1 before importing triton"
    os.environ["TRITON_INTERPRET"] = "1"
    os.environ["TORCHDYNAMO_DISABLE"] = "1"
    if REF_LONGCAT not in sys.path:
        sys.path.insert(0, REF_LONGCAT)
    for name in ("longcat_video", "longcat_video.context_parallel", "longcat_video.block_sparse_attention"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_LONGCAT, *name.split("."))]
            sys.modules[name] = m
    name = "longcat_video.block_sparse_attention.bsa_interface"
    if name in sys.modules and getattr(sys.modules[name], "flash_attn_bsa_3d", None) is None:
        del sys.modules[name]                                   # the stub of load_longcat_dit_module(bsa=False)
    return importlib.import_module(name)


def load_longcat_dit_module(bsa: bool = False):
    """longcat_video.modules.longcat_video_dit, unmodified, importable on the CPU: diffusers stand-in, a stub for the
    Triton block-sparse package (only reached by the 720p refine pass; ``bsa=True`` loads the real one under Triton's
    interpreter), SDPA-backed xformers, context-parallel size 1."""
    assert os.path.isdir(REF_LONGCAT)
    if bsa:
        real = load_longcat_bsa_module()
        att = sys.modules.get("longcat_video.modules.attention")
        if att is not None:                                     # imported earlier against the stub: rebind the one name it took
            att.flash_attn_bsa_3d = real.flash_attn_bsa_3d
    install_diffusers_shim()
    _fake_xformers()
    if REF_LONGCAT not in sys.path:
        sys.path.insert(0, REF_LONGCAT)
    for name in ("longcat_video", "longcat_video.modules", "longcat_video.context_parallel", "longcat_video.block_sparse_attention"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_LONGCAT, *name.split("."))]
            sys.modules[name] = m
    if "longcat_video.block_sparse_attention.bsa_interface" not in sys.modules:
        stub = types.ModuleType("longcat_video.block_sparse_attention.bsa_interface")
        stub.flash_attn_bsa_3d = None
        sys.modules["longcat_video.block_sparse_attention.bsa_interface"] = stub
    cpu = importlib.import_module("longcat_video.context_parallel.context_parallel_util")
    cpu.cp_size = 1
    return importlib.import_module("longcat_video.modules.longcat_video_dit")


def load_longcat_vae_module():
    """longcat_video/modules/autoencoder_kl_wan.py - the vendored copy of diffusers' ``AutoencoderKLWan`` (the class
    infer_worldforge.py:185-189 loads) - over the shim, for pinning the diffusers <-> WanVAE_ parameter-name map."""
    install_diffusers_shim()
    import torch.nn as nn

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    ld = sys.modules.get("diffusers.loaders") or mod("diffusers.loaders")
    ld.FromOriginalModelMixin = type("FromOriginalModelMixin", (), {})
    au = mod("diffusers.utils.accelerate_utils")
    au.apply_forward_hook = lambda f: f
    act = mod("diffusers.models.activations")
    act.get_activation = lambda name: {"silu": nn.SiLU(), "swish": nn.SiLU(), "gelu": nn.GELU()}[name]
    class _AttrOut(dict):
        __getattr__ = dict.__getitem__

    mo = mod("diffusers.models.modeling_outputs")
    mo.AutoencoderKLOutput = _AttrOut
    ae = mod("diffusers.models.autoencoders"); av = mod("diffusers.models.autoencoders.vae")
    av.DecoderOutput = _AttrOut

    class _Diag:
        def __init__(self, parameters):
            self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)

        def mode(self):
            return self.mean
    av.DiagonalGaussianDistribution = _Diag
    ae.vae = av
    path = os.path.join(REF_ROOT, "longcat_for_worldforge", "longcat_video", "modules", "autoencoder_kl_wan.py")
    spec = importlib.util.spec_from_file_location("_wf_ref_autoencoder_kl_wan", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_longcat_scheduler_module():
    assert os.path.isdir(REF_LONGCAT)
    install_diffusers_shim()
    spec = importlib.util.spec_from_file_location(
        "wf_ref_longcat_scheduling", os.path.join(REF_LONGCAT, "longcat_video", "modules", "scheduling_flow_match_euler_discrete.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


# ------------------------------------------------------------------------------------------------------------------
# The reference PIPELINES: WanImageToVideoPipeline.__call__ and LongCatVideoPipeline.generate_i2v run
# UNMODIFIED over oracle objects, so that oracle/pipeline.py and oracle/longcat_sched.py - restatements of those two
# loops - are pinned against the loops themselves and not only against the schedulers they drive.
# ------------------------------------------------------------------------------------------------------------------

class _Progress:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a, **k):
        pass


class _DiffusionPipeline:
    """Stand-in for diffusers.DiffusionPipeline with the five members pipeline_wan_i2v_clean.py uses:
    register_modules (:152), _execution_device (:484), progress_bar (:562), maybe_free_model_hooks (:748) and the
    plain constructor."""
    exec_device = torch.device("cpu")

    def __init__(self, *a, **k):
        pass

    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def _execution_device(self):
        return self.exec_device

    def progress_bar(self, iterable=None, total=None):
        return _Progress()

    def maybe_free_model_hooks(self):
        pass


class _VideoProcessor:
    """Stand-in for diffusers.video_processor.VideoProcessor.  ``preprocess`` is only defined here for what the pinning
    runs feed it - a [B,3,H,W] float tensor already in [-1,1] at the target size, which diffusers' VaeImageProcessor
    returns unchanged (it skips its [0,1]->[-1,1] normalisation when the tensor has negative values) - and raises for
    anything else instead of guessing."""

    def __init__(self, vae_scale_factor=8, **kw):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height=None, width=None):
        if not (isinstance(image, torch.Tensor) and image.dim() == 4 and image.shape[-2:] == (height, width)
                and float(image.min()) < 0 and float(image.abs().max()) <= 1):
            raise NotImplementedError("shim VideoProcessor: pass a [B,3,H,W] tensor in [-1,1] of the target size")
        return image

    def preprocess_video(self, video, height=None, width=None):
        """[B,3,F,H,W] float tensor already in [-1,1] at the target size (what diffusers returns for such an input)."""
        if not (isinstance(video, torch.Tensor) and video.dim() == 5 and video.shape[-2:] == (height, width)
                and float(video.min()) < 0 and float(video.abs().max()) <= 1):
            raise NotImplementedError("shim VideoProcessor: pass a [B,3,F,H,W] tensor in [-1,1] of the target size")
        return video

    def postprocess_video(self, video, output_type="np"):
        raise NotImplementedError("shim VideoProcessor: run the pinned pipelines with output_type='latent'")


def _randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor for a single CPU generator: drawn on the generator's device, then moved."""
    return torch.randn(shape, generator=generator, dtype=dtype).to(device)


def install_pipeline_shim() -> None:
    """The remaining ``diffusers`` names the reference pipelines import (pipeline_wan_i2v_clean.py:22-31,
    pipeline_longcat_video.py:11-12)."""
    install_diffusers_shim()
    d = sys.modules["diffusers"]
    if not getattr(d, "_wf_shim", False):
        return
    def mod(name):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
        return sys.modules[name]
    cb = mod("diffusers.callbacks")
    cb.PipelineCallback = type("PipelineCallback", (), {})
    cb.MultiPipelineCallbacks = type("MultiPipelineCallbacks", (), {})
    mod("diffusers.image_processor").PipelineImageInput = object
    mod("diffusers.loaders").WanLoraLoaderMixin = type("WanLoraLoaderMixin", (), {})
    mm = sys.modules["diffusers.models"]
    mm.AutoencoderKLWan = type("AutoencoderKLWan", (), {})
    mm.WanTransformer3DModel = type("WanTransformer3DModel", (), {})
    sys.modules["diffusers.schedulers"].FlowMatchEulerDiscreteScheduler = type("FlowMatchEulerDiscreteScheduler", (), {})
    ut = sys.modules["diffusers.utils"]
    ut.is_ftfy_available = lambda: False
    ut.is_torch_xla_available = lambda: False
    ut.replace_example_docstring = lambda doc: (lambda fn: fn)
    tu = mod("diffusers.utils.torch_utils"); tu.randn_tensor = _randn_tensor
    ut.torch_utils = tu
    mod("diffusers.video_processor").VideoProcessor = _VideoProcessor
    pp = mod("diffusers.pipelines"); pu = mod("diffusers.pipelines.pipeline_utils")
    pu.DiffusionPipeline = _DiffusionPipeline
    pw = mod("diffusers.pipelines.wan"); po = mod("diffusers.pipelines.wan.pipeline_output")
    po.WanPipelineOutput = lambda frames: types.SimpleNamespace(frames=frames)
    pp.pipeline_utils, pp.wan, pw.pipeline_output = pu, pw, po
    for name in ("callbacks", "image_processor", "loaders", "video_processor", "pipelines"):
        setattr(d, name, sys.modules["diffusers." + name])


def load_wan_pipeline_module():
    """utils/pipeline_wan_i2v_clean.py, unmodified (its ``transformers`` imports resolve to the installed package)."""
    assert available()
    install_pipeline_shim()
    spec = importlib.util.spec_from_file_location(
        "wf_ref_pipeline_wan_i2v", os.path.join(REF_WAN, "utils", "pipeline_wan_i2v_clean.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_longcat_pipeline_module():
    """longcat_video/pipeline_longcat_video.py, unmodified: ``LongCatVideoPipeline`` (a plain class).  Its own package
    imports resolve to the reference tree (DiT / scheduler / VAE modules over the shims above); ``loguru`` and ``ftfy`` are
    not installed and get no-op stand-ins (logging and prompt cleaning only)."""
    assert os.path.isdir(REF_LONGCAT)
    install_pipeline_shim()
    load_longcat_vae_module()                       # registers the diffusers names autoencoder_kl_wan.py imports
    _fake_xformers()
    if "loguru" not in sys.modules:
        lg = types.ModuleType("loguru")
        lg.logger = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None, error=lambda *a, **k: None,
                                          debug=lambda *a, **k: None)
        sys.modules["loguru"] = lg
    if "ftfy" not in sys.modules:
        f = types.ModuleType("ftfy"); f.fix_text = lambda t: t
        sys.modules["ftfy"] = f
    if REF_LONGCAT not in sys.path:
        sys.path.insert(0, REF_LONGCAT)
    for name in [m for m in sys.modules if m == "longcat_video" or m.startswith("longcat_video.")]:
        del sys.modules[name]
    return importlib.import_module("longcat_video.pipeline_longcat_video")


class FixedTextEncoder:
    """``text_encoder`` stand-in for LongCatVideoPipeline.encode_prompt (pipeline_longcat_video.py:90-189), with its
    ``tokenizer`` twin: the prompt string selects one of the given (embedding [L, C], valid length) pairs; everything the
    pipeline does with them afterwards (masking, the [negative, positive] batch) is the reference's own code."""

    def __init__(self, table, max_len: int, dim: int):
        self.table, self.max_len = table, max_len
        self.dtype = torch.bfloat16
        self.config = types.SimpleNamespace(d_model=dim)
        self._order = list(table)
        enc = self

        class _Tokenizer:
            def __call__(self, prompt, **kw):
                ids = torch.zeros(len(prompt), enc.max_len, dtype=torch.long)
                mask = torch.zeros(len(prompt), enc.max_len, dtype=torch.long)
                for i, p in enumerate(prompt):
                    ids[i, 0] = enc._order.index(p)
                    mask[i, :enc.table[p][1]] = 1
                return types.SimpleNamespace(input_ids=ids, attention_mask=mask)
        self.tokenizer = _Tokenizer()

    def __call__(self, ids, mask):
        hs = torch.stack([self.table[self._order[int(i[0])]][0] for i in ids])
        return types.SimpleNamespace(last_hidden_state=hs)


class FixedImageEncoder:
    """``image_processor`` + ``image_encoder`` stand-ins for encode_image (pipeline_wan_i2v_clean.py:205-209): the
    reference refuses ``image`` together with ``image_embeds`` (:362-363) but needs ``image`` for prepare_latents, so the
    CLIP embedding enters through these two objects; the second-to-last hidden state is the tensor given here."""

    def __init__(self, image_embeds):
        self.embeds = image_embeds

    def processor(self, images=None, return_tensors="pt"):
        class _Batch(dict):
            def to(self, device):
                return self
        return _Batch(pixel_values=images)

    def encoder(self, pixel_values=None, output_hidden_states=True):
        return types.SimpleNamespace(hidden_states=[None, self.embeds, None])


# ---- "a parity harness must assert these fallbacks were not taken" ------------------------------------------------
class no_silent_fallbacks:
    """The reference schedulers swallow failures and carry on with a different computation (fuse_latents returns the unfused
    prediction, the selectors drop from Farneback flow to temporal differences / gradients, a failed metric scores 0.0 - all
    inside bare ``except`` blocks: scheduling_unipc_multistep_clean.py:111, :225, :390, :477, :492, :606, :1286, :1414, :1419;
    scheduling_flow_match_euler_discrete.py:142, :161, :242, :301, :322, :885, :1215, :1224).  A fixture generated through one
    of them would pin the wrong computation.  Inside this context every such path leaves a trace in ``events``; the golden
    generators assert the list stays empty."""

    FALLBACK_METHODS = ("_extract_multiscale_motion", "_extract_temporal_difference")

    def __init__(self, selector_cls):
        self.cls, self.events, self._undo = selector_cls, [], []

    def _patch(self, obj, name, fn):
        self._undo.append((obj, name, getattr(obj, name)))
        setattr(obj, name, fn)

    def __enter__(self):
        import logging
        import cv2
        ev = self.events
        for name in self.FALLBACK_METHODS:
            if hasattr(self.cls, name):
                orig = getattr(self.cls, name)
                def spy(self_, *a, _orig=orig, _name=name, **k):
                    ev.append(("fallback method", _name))
                    return _orig(self_, *a, **k)
                self._patch(self.cls, name, spy)
        if hasattr(self.cls, "_compute_flow_metrics"):
            orig = self.cls._compute_flow_metrics
            def metrics(self_, *a, _orig=orig, **k):
                r = _orig(self_, *a, **k)
                if r == 0.0:                       # the value its ``except`` returns; a genuine score of exactly 0 is not expected
                    ev.append(("metric 0.0", None))
                return r
            self._patch(self.cls, "_compute_flow_metrics", metrics)
        orig_fb = cv2.calcOpticalFlowFarneback
        def farneback(*a, **k):
            try:
                return orig_fb(*a, **k)
            except Exception as ex:
                ev.append(("cv2.calcOpticalFlowFarneback raised", repr(ex)))
                raise
        self._patch(cv2, "calcOpticalFlowFarneback", farneback)

        class Handler(logging.Handler):
            def emit(self_, record):
                if record.levelno >= logging.WARNING:
                    ev.append(("log", record.getMessage()))
        self._handler = Handler()
        logging.getLogger().addHandler(self._handler)
        return self.events

    def __exit__(self, *a):
        import logging
        logging.getLogger().removeHandler(self._handler)
        for obj, name, orig in reversed(self._undo):
            setattr(obj, name, orig)
        self._undo = []
        return False
