"""On-disk artefacts exchanged with the reference's scripts, byte-compatible with what they read and write.

  means_flattened_<uid>, stds_flattened_<uid>     torch.save of 1-D fp32 tensors     sensitivity.py:219-222, read at main_VI_HMC.py:76-77
  gradient_indices_<uid>.npy                      sorted int64 indices               sensitivity.py:231-233, read at main_VI_HMC.py:79,355
  mean_importance_<uid>.npy                       fp32 scores                        sensitivity.py:234
  hmc_params_<uid>.npy                            (S, d) fp32, one row per draw      main_VI_HMC.py:381, read back at :418 and sliced [cfg.burn:]

The engine samples many chains at once; one file per chain keeps the reference's (S, d) layout, so its post-processing
(`params_hmc[cfg.burn:]`, predict_model over the rows) runs unchanged on any of them.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch


def save_vi_artifacts(directory: str, uid: str, means: torch.Tensor, stds: torch.Tensor, grad_ind: Optional[Sequence[int]] = None,
                      importance: Optional[np.ndarray] = None) -> None:
    os.makedirs(directory, exist_ok=True)
    torch.save(means.detach().float().cpu().reshape(-1), os.path.join(directory, f"means_flattened_{uid}"))
    torch.save(stds.detach().float().cpu().reshape(-1), os.path.join(directory, f"stds_flattened_{uid}"))
    if grad_ind is not None:
        np.save(os.path.join(directory, f"gradient_indices_{uid}.npy"), np.sort(np.asarray(grad_ind, dtype=np.int64)))
    if importance is not None:
        np.save(os.path.join(directory, f"mean_importance_{uid}.npy"), np.asarray(importance))


def load_vi_artifacts(directory: str, uid: str) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    """(means [D], stds [D], gradient_indices [d]) as main_VI_HMC.py:76-79 loads them."""
    means = torch.load(os.path.join(directory, f"means_flattened_{uid}"), map_location="cpu")
    stds = torch.load(os.path.join(directory, f"stds_flattened_{uid}"), map_location="cpu")
    ind = np.load(os.path.join(directory, f"gradient_indices_{uid}.npy"), allow_pickle=True)
    return means, stds, ind


def save_hmc_params(out_dir: str, uid: str, samples: Union[torch.Tensor, List[torch.Tensor]]) -> List[str]:
    """Write hmc_params_<uid>.npy.  `samples`: the list of 1-D tensors `sample()` returns for one chain (saved exactly as
    the reference's `np.save(path, params_hmc)` does), or a [S, C, d] tensor -- then one file per chain, <uid>_c<k>."""
    os.makedirs(out_dir, exist_ok=True)
    if isinstance(samples, (list, tuple)):
        arr = np.stack([np.asarray(t.detach().cpu(), dtype=np.float32) for t in samples])
        path = os.path.join(out_dir, f"hmc_params_{uid}.npy")
        np.save(path, arr)
        return [path]
    arr = samples.detach().cpu().numpy().astype(np.float32, copy=False)
    if arr.ndim == 2:
        path = os.path.join(out_dir, f"hmc_params_{uid}.npy")
        np.save(path, arr)
        return [path]
    paths = []
    for c in range(arr.shape[1]):
        path = os.path.join(out_dir, f"hmc_params_{uid}_c{c}.npy")
        np.save(path, np.ascontiguousarray(arr[:, c, :]))
        paths.append(path)
    return paths


def load_hmc_params(path: str, burn: int = 0) -> torch.Tensor:
    """`torch.tensor(np.load(path))[burn:]` -- main_VI_HMC.py:418-420."""
    return torch.tensor(np.load(path))[burn:]
