import pytest

pytestmark = pytest.mark.gpu


def test_smoke_case():
    import smoke_case
    smoke_case.run(verbose=True)
