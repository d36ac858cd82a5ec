"""CPU checks of the source-compatibility layer: the `pybnesian` alias package (the reference is one flat module of that
name, /root/reference/pybnesian/lib.cpp:22-51), the Arrow C-Data entry points of the DataFrame wrapper
(dataset/dataset.hpp:2088-2143) and the device selection of the default context."""
import numpy as np
import pyarrow as pa
import pytest


def test_pybnesian_alias_reexports_the_path():
    import pybnesian as pbn
    import pybnesian_b200 as impl
    for name in ("KDE", "CKDE", "ProductKDE", "LinearGaussianCPD", "CKDEType", "LinearGaussianCPDType", "BandwidthSelector",
                 "NormalReferenceRule", "ScottsBandwidth", "UCV", "UCVScorer", "CVLikelihood", "HoldoutLikelihood",
                 "ValidatedLikelihood", "CrossValidation", "HoldOut", "GreedyHillClimbing", "hc", "SemiparametricBN",
                 "GaussianNetwork", "KDENetwork", "ArcOperatorSet", "ChangeNodeTypeSet", "OperatorPool", "OperatorTabuSet",
                 "AddArc", "RemoveArc", "FlipArc", "ChangeNodeType", "BIC", "MLE", "load", "SingularCovarianceData",
                 "BayesianNetwork", "BayesianNetworkType", "UnknownFactorType"):
        assert getattr(pbn, name) is getattr(impl, name), name
    assert issubclass(pbn.SingularCovarianceData, ValueError)
    with pytest.raises(AttributeError, match="outside the scope"):
        pbn.ConditionalGaussianNetwork


def test_dataframe_takes_arrow_c_data_capsules():
    from pybnesian_b200.dataset import DataFrame, _to_record_batch
    rb = pa.RecordBatch.from_arrays([pa.array(np.arange(5.0)), pa.array(np.ones(5, dtype=np.float32))], names=["a", "b"])
    assert _to_record_batch(rb.__arrow_c_array__()).equals(rb)          # (schema, array) PyCapsule pair

    class Exporter:      # any object speaking the PyCapsule protocol (polars, nanoarrow, duckdb results, ...)
        def __arrow_c_array__(self, requested_schema=None):
            return rb.__arrow_c_array__(requested_schema)

    class Stream:
        def __arrow_c_stream__(self, requested_schema=None):
            return pa.Table.from_batches([rb, rb]).__arrow_c_stream__(requested_schema)

    assert DataFrame(Exporter()).rb.equals(rb)
    assert DataFrame(Stream()).num_rows == 10
    f = DataFrame(Exporter())
    assert f.columns == ["a", "b"] and f.same_type(["a"]) == pa.float64()
    with pytest.raises(ValueError, match="different data types"):
        f.same_type(["a", "b"])
    with pytest.raises(TypeError):
        DataFrame(object())


def test_default_devices_from_the_environment(monkeypatch):
    from pybnesian_b200 import _lib
    for var in ("PBN_CUDA_DEVICES", "PBN_CUDA_DEVICE", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert _lib._default_devices() == [3]                 # torchrun: one process per GPU
    monkeypatch.setenv("PBN_CUDA_DEVICE", "1")
    assert _lib._default_devices() == [1]                 # explicit single device wins over LOCAL_RANK
    monkeypatch.setenv("PBN_CUDA_DEVICES", "0, 2,5")
    assert _lib._default_devices() == [0, 2, 5]           # explicit list wins over everything: one multi-device context
