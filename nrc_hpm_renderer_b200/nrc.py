"""Python mirror of ``en::NeuralRadianceCache`` (reference include/engine/graphics/NeuralRadianceCache.hpp:10-64,
src/NeuralRadianceCache.cu:11-178) on top of the C ABI.  Same method names and argument meaning; device buffers are
torch CUDA tensors (torch is only the allocator / stream provider here), host buffers are numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _lib
from .config import AppConfig

MASTER, WORKING, EMA, GRAD, ADAM_M, ADAM_V, STEPS = range(7)
SNAPSHOT = 2          # `use_ema` value of inference(): the parameters captured by snapshot_params()


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dptr(t) -> int:
    """device pointer of a torch CUDA tensor (or a raw integer address)"""
    if isinstance(t, int):
        return t
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


class NeuralRadianceCache:
    """``NeuralRadianceCache(appConfig)``: builds the model from the same JSON the reference hands to tiny-cuda-nn."""

    def __init__(self, app_config: AppConfig | None = None, *, config_json: dict | str | None = None, seed: int = 1337,
                 oneblob_soa_bug: bool = True):
        if config_json is None:
            app_config = app_config or AppConfig.default()
            cfg = app_config.model_json()
            cfg["infer_batch_size"] = app_config.infer_batch_size
            cfg["train_batch_size"] = app_config.train_batch_size
            cfg["train_batch_count"] = app_config.train_batch_count
            cfg["compat"] = {"oneblob_soa_bug": bool(oneblob_soa_bug)}
            config_json = cfg
        text = config_json if isinstance(config_json, str) else json.dumps(config_json)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().nrc_create(text.encode(), seed, C.byref(self._h)))
        self._keep = []

    # ---- en::NeuralRadianceCache surface
    def Init(self, infer_count, d_infer_input, d_infer_output, d_train_input, d_train_target, start_semaphore=None,
             finished_semaphore=None, stream=None):
        self._keep = [d_infer_input, d_infer_output, d_train_input, d_train_target]
        _lib.check(_lib.lib().nrc_init(self._h, infer_count, _dptr(d_infer_input), _dptr(d_infer_output), _dptr(d_train_input),
                                       _dptr(d_train_target), start_semaphore, finished_semaphore, stream))

    def InferAndTrain(self, infer_filter: np.ndarray | None, train: bool):
        f = None if infer_filter is None else np.ascontiguousarray(infer_filter, np.uint32).ctypes.data_as(C.POINTER(C.c_uint32))
        _lib.check(_lib.lib().nrc_infer_and_train(self._h, f, int(train)))

    def Inference(self, infer_filter: np.ndarray | None = None):
        f = None if infer_filter is None else np.ascontiguousarray(infer_filter, np.uint32).ctypes.data_as(C.POINTER(C.c_uint32))
        _lib.check(_lib.lib().nrc_inference(self._h, f))

    def Train(self):
        _lib.check(_lib.lib().nrc_train(self._h))

    def Destroy(self):
        if self._h:
            _lib.check(_lib.lib().nrc_destroy(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass

    def GetLoss(self) -> float:
        v = C.c_float()
        _lib.check(_lib.lib().nrc_get_loss(self._h, C.byref(v)))
        return float(v.value)

    def GetInferBatchCount(self) -> int: return int(_lib.lib().nrc_get_infer_batch_count(self._h))
    def GetTrainBatchCount(self) -> int: return int(_lib.lib().nrc_get_train_batch_count(self._h))
    def GetInferBatchSize(self) -> int: return int(_lib.lib().nrc_get_infer_batch_size(self._h))
    def GetTrainBatchSize(self) -> int: return int(_lib.lib().nrc_get_train_batch_size(self._h))

    # ---- tcnn-level surface
    @property
    def n_params(self) -> int: return int(_lib.lib().nrc_n_params(self._h))
    @property
    def n_mlp_params(self) -> int: return int(_lib.lib().nrc_n_mlp_params(self._h))
    @property
    def input_width(self) -> int: return int(_lib.lib().nrc_input_width(self._h))

    def get_params(self, which=MASTER) -> np.ndarray:
        out = np.empty(self.n_params, np.float32)
        _lib.check(_lib.lib().nrc_get_params(self._h, which, _fp(out)))
        return out

    def set_params(self, master: np.ndarray):
        _lib.check(_lib.lib().nrc_set_params_fp32(self._h, _fp(np.ascontiguousarray(master, np.float32))))

    def set_ema(self, ema: np.ndarray):
        _lib.check(_lib.lib().nrc_set_ema(self._h, _fp(np.ascontiguousarray(ema, np.float32))))

    def encode(self, d_in, n, d_out_half, use_ema=False, stream=None):
        _lib.check(_lib.lib().nrc_encode_batch(self._h, _dptr(d_in), n, int(use_ema), _dptr(d_out_half), stream))

    def inference(self, d_in, d_out, n, use_ema=True, stream=None):
        _lib.check(_lib.lib().nrc_inference_batch(self._h, _dptr(d_in), _dptr(d_out), n, int(use_ema), stream))

    def snapshot_params(self, use_ema=True, stream=None):
        """capture the parameters inference(..., use_ema=SNAPSHOT) evaluates (stream-ordered device-to-device copy)"""
        _lib.check(_lib.lib().nrc_snapshot_params(self._h, int(use_ema), stream))

    PEER_HANDLE_BYTES = 192

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(self.PEER_HANDLE_BYTES)
        _lib.check(_lib.lib().nrc_peer_export(self._h, buf))
        return buf.raw

    def peer_setup(self, rank: int, world: int, all_handles: bytes):
        assert len(all_handles) == world * self.PEER_HANDLE_BYTES
        _lib.check(_lib.lib().nrc_peer_setup(self._h, rank, world, C.create_string_buffer(all_handles, len(all_handles))))

    def peer_exchange(self, stream=None):
        _lib.check(_lib.lib().nrc_peer_exchange(self._h, stream))

    def set_inference_cta_limit(self, max_ctas: int):
        _lib.check(_lib.lib().nrc_set_inference_cta_limit(self._h, int(max_ctas)))

    def inference_indexed(self, d_in, d_out, d_indices, d_count, max_n, use_ema=True, stream=None):
        _lib.check(_lib.lib().nrc_inference_indexed(self._h, _dptr(d_in), _dptr(d_out), _dptr(d_indices),
                                                    None if d_count is None else _dptr(d_count), max_n, int(use_ema), stream))

    def training_step(self, d_in, d_target, batch, run_optimizer=True, stream=None):
        _lib.check(_lib.lib().nrc_training_step(self._h, _dptr(d_in), _dptr(d_target), batch, int(run_optimizer), stream))

    def optimizer_step(self, stream=None):
        _lib.check(_lib.lib().nrc_optimizer_step(self._h, stream))

    def gradient_buffers(self):
        """(device pointer of the fp32 MLP gradient [n_mlp], device pointer of the fp16 encoding gradient or None)"""
        a, b = C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib().nrc_gradient_buffers(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_step_tensor(self, which: int, batch: int) -> np.ndarray:
        width = 16 if which < 2 else self.input_width
        out = np.empty((batch, width), np.float32)
        _lib.check(_lib.lib().nrc_last_step_tensor(self._h, which, _fp(out)))
        return out

    # ---- host-buffer entry points (H2D / D2H inside the call)
    def inference_host(self, rec: np.ndarray, use_ema=True, out: np.ndarray | None = None) -> np.ndarray:
        rec = np.ascontiguousarray(rec, np.float32)
        if out is None:
            out = np.empty((len(rec), 3), np.float32)
        _lib.check(_lib.lib().nrc_inference_host(self._h, _fp(rec), _fp(out), len(rec), int(use_ema)))
        return out

    def training_step_host(self, rec: np.ndarray, target: np.ndarray) -> float:
        rec = np.ascontiguousarray(rec, np.float32); target = np.ascontiguousarray(target, np.float32)
        v = C.c_float()
        _lib.check(_lib.lib().nrc_training_step_host(self._h, _fp(rec), _fp(target), len(rec), C.byref(v)))
        return float(v.value)

    def infer_and_train_host(self, rec: np.ndarray, out: np.ndarray, train_rec: np.ndarray | None, train_target: np.ndarray | None,
                             batch: int = 0, use_ema=True) -> float:
        """``InferAndTrain`` on host buffers in one call (pipelined copies, one host wait); returns the last batch's loss"""
        n = 0 if rec is None else len(rec)
        nb = 0 if train_rec is None or batch == 0 else len(train_rec) // batch
        v = C.c_float()
        _lib.check(_lib.lib().nrc_infer_and_train_host(self._h, _fp(rec) if n else None, _fp(out) if n else None, n,
                                                       _fp(train_rec) if nb else None, _fp(train_target) if nb else None, batch, nb, int(use_ema), C.byref(v)))
        return float(v.value)
