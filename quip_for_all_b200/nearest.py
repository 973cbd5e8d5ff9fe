"""Binding of `quipb200_e8p_quantize` (include/quip_b200.h): the quantise-time nearest-codeword search of the E8P12
family on the GPU (reference: codebook/e8p12.py:125-134, codebook/e8p12_rvq4.py:32-46).  Called by
`E8P12_codebook.quantize` / `E8P12RVQ4B_codebook.quantize` for CUDA fp32 inputs -- the case LDLQ produces
(quant.py:128-129, with the default `use_fp64=False`)."""
import ctypes

import torch

from ._native import check, lib


def native_ok(X: torch.Tensor) -> bool:
    return X.is_cuda and X.dtype == torch.float32 and X.dim() >= 1 and X.shape[-1] == 8 and X.numel() > 0


def e8p_quantize(X: torch.Tensor, grid_packed_abs: torch.Tensor, n_stages: int = 1, resid_scale: float = 1.0,
                 resid_grid: torch.Tensor = None):
    """X: CUDA fp32 [..., 8].  Returns (vals fp32 [..., 8], idx int64 [...]) exactly as the reference's `quantize`.
    `resid_grid` (fp32 [256, 8]): the second search runs against this table (E8P12RVQ3B's e81b grid)."""
    if not native_ok(X):
        raise ValueError("e8p_quantize: CUDA float32 [..., 8] input required")
    if grid_packed_abs.device != X.device or grid_packed_abs.dtype != torch.int64 or grid_packed_abs.numel() != 256:
        raise ValueError("e8p_quantize: grid_packed_abs must be the int64[256] table on the input's device")
    x = X.reshape(-1, 8).contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    m = x.shape[0]
    L = lib()
    vals = torch.empty_like(x)
    idx = torch.empty(m, dtype=torch.int64, device=x.device)
    ws_bytes = L.quipb200_e8p_quantize_workspace_bytes(m)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if resid_grid is not None:
            rg = resid_grid.to(device=x.device, dtype=torch.float32).contiguous()
            if rg.shape != (256, 8):
                raise ValueError("e8p_quantize: resid_grid must be [256, 8]")
            check(L.quipb200_e8prvq3_quantize(x.data_ptr(), m, grid_packed_abs.data_ptr(), rg.data_ptr(), float(resid_scale),
                                              vals.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws_bytes, st), "e8prvq3_quantize")
        else:
            check(L.quipb200_e8p_quantize(x.data_ptr(), m, grid_packed_abs.data_ptr(), int(n_stages), float(resid_scale),
                                          vals.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws_bytes, st), "e8p_quantize")
    return vals.view(*X.shape), idx.view(*X.shape[:-1])
