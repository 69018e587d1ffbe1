"""File-name and vector helpers the hot path needs, with the reference's names and results.

Only the pieces ``cu_knn`` / ``cu_knn_ds`` / ``merge`` use are here
(neighborhoodwatch/nw_utils.py:26-30, 56-59, 94-107, 157-164); the dataset / CLI helpers of
that module are outside the hot path.  Unlike the reference module
this one imports neither ``datasets`` nor the embedding SDKs.
"""
from __future__ import annotations

import os

import numpy as np

# model_generator.py:38 -- the one model whose stored dimension may be 1/8 of the nominal one
VOYAGE_3_LARGE = "voyage-3-large"


def get_full_filename(data_dir: str, filename: str) -> str:
    """nw_utils.py:26-30: prefix ``data_dir`` unless the name already starts with it."""
    return filename if filename.startswith(data_dir) else f"{data_dir}/{filename}"


def get_partial_indices_filename(knn_model_data_homedir: str, partial_set_cnt: int) -> str:
    """nw_utils.py:94-99: ``partial/indices{i}.parquet``; ``-1`` names the merged result."""
    stem = "final_indices" if partial_set_cnt == -1 else f"indices{partial_set_cnt}"
    return f"{knn_model_data_homedir}/partial/{stem}.parquet"


def get_partial_distances_filename(knn_model_data_homedir: str, partial_set_cnt: int) -> str:
    """nw_utils.py:102-107."""
    stem = "final_distances" if partial_set_cnt == -1 else f"distances{partial_set_cnt}"
    return f"{knn_model_data_homedir}/partial/{stem}.parquet"


def ensure_partial_dir(data_dir: str) -> str:
    """The reference creates ``<data_dir>/partial`` in setup_model_output_folder (nw_utils.py:62-73)."""
    path = os.path.join(data_dir, "partial")
    os.makedirs(path, exist_ok=True)
    return path


def is_zero_embedding(embedding) -> bool:
    return not np.any(embedding)


def normalize_vector(vector):
    """nw_utils.py:56-59: L2-normalise, refusing the zero vector, result float32."""
    assert not is_zero_embedding(vector), "Zero vector found!"
    return (vector / np.linalg.norm(vector)).astype(np.float32)


def output_dimension_validity_check(model_name, given_dimension: int, actual_dimension: int) -> bool:
    """nw_utils.py:157-164: column count must equal ``dimensions`` (Voyage binary: 1/8)."""
    ok = actual_dimension == given_dimension
    if model_name == VOYAGE_3_LARGE:
        ok = ok or given_dimension == 8 * actual_dimension
    return ok
