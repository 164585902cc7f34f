"""Configuration contract of the train_colvars step.

Field names and defaults follow the reference's ``yaml_schemas/train_colvars.py:162-249`` for
the fields the CV hot path reads; everything else the reference accepts (figures, bias,
architecture details, ...) is tolerated through ``extra='allow'`` so reference YAMLs stay valid.
Backend knobs live in the optional ``backend`` block.
"""
from typing import Dict, List, Literal, Optional

from pydantic import BaseModel, ConfigDict


class InputColvars(BaseModel):
    model_config = ConfigDict(extra="allow")
    start: int = 0
    stop: Optional[int] = None
    stride: int = 1


class Backend(BaseModel):
    """B200 backend options (not in the reference)."""
    model_config = ConfigDict(extra="allow")
    # covariance contraction engine: auto (= tc_i8x3, the exact integer tensor-core engine), or one of
    # the float tcgen05 engines / the CUDA-core FP32 engine explicitly
    cov_engine: Literal["auto", "tc_i8x3", "tc_3xf16", "tc_3xtf32", "tc_1xtf32", "simt_f32"] = "auto"
    # hTICA: accumulate the full F x F Gram in one pass when F <= this, else block-diagonal + 2nd pass
    htica_full_gram_max_features: int = 2048
    # CUDA device index (None = current device / LOCAL_RANK)
    device: Optional[int] = None


class CommonCollectiveVariable(BaseModel):
    model_config = ConfigDict(extra="allow")
    dimension: int = 2
    lag_time: int = 1
    tica_regularization: float = 1.0e-06
    features_normalization: Optional[Literal["mean_std", "min_max_range1", "min_max_range2"]] = None
    input_colvars: InputColvars = InputColvars()
    architecture: Dict = {}
    training: Dict = {}
    num_subspaces: int = 10
    subspaces_dimension: int = 5
    backend: Backend = Backend()


class TrainColvarsSchema(BaseModel):
    model_config = ConfigDict(extra="allow")
    cvs: List[Literal["pca", "ae", "tica", "htica", "deep_tica", "vae", "umap"]] = ["pca", "tica", "htica"]
    common: CommonCollectiveVariable = CommonCollectiveVariable()
    figures: Dict = {}
