"""Host-side pieces of the CLI front-end (no GPU): vector parsing / normalisation (src/main.rs:259-281)."""
import numpy as np
import pytest

from chronomind_b200 import cli


def test_parse_vector_accepts_json_and_csv():
    assert np.array_equal(cli.parse_vector(" [0.1, 0.2, 3] "), np.array([0.1, 0.2, 3.0], np.float32))
    assert np.array_equal(cli.parse_vector("0.1, 0.2 ,3"), np.array([0.1, 0.2, 3.0], np.float32))
    with pytest.raises(SystemExit):
        cli.parse_vector("0.1,abc")


def test_normalize_vector():
    v = cli.normalize_vector(np.array([3.0, 4.0], np.float32))
    assert np.allclose(v, [0.6, 0.8])
    z = np.zeros(3, np.float32)
    assert np.array_equal(cli.normalize_vector(z), z)          # norm == 0: left alone
