import pytest


@pytest.fixture(autouse=True)
def isolated_home(tmp_path, monkeypatch):
    """Every test gets its own ~/.prime (the reference isolates per xdist worker: prime-sandboxes/tests/conftest.py:12-28)."""
    home = tmp_path / "home"
    home.mkdir()
    monkeypatch.setenv("HOME", str(home))
    monkeypatch.setattr("pathlib.Path.home", lambda: home)
    for k in ("PRIME_API_KEY", "PRIME_TEAM_ID", "PRIME_USER_ID", "PRIME_API_BASE_URL", "PRIME_BASE_URL", "PRIME_CONTEXT",
              "PRIME_FRONTEND_URL", "PRIME_INFERENCE_URL", "PRIME_SSH_KEY_PATH"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("PRIME_DISABLE_VERSION_CHECK", "1")
    monkeypatch.setenv("COLUMNS", "200")
    return home


@pytest.fixture
def anyio_backend():
    return "asyncio"


class FakeAPI:
    """Stand-in for APIClient in command tests: routes by (METHOD, path) → payload or callable, records every call."""

    def __init__(self, routes=None, api_key="k-test"):
        from prime_b200.platform.core import Config

        self.routes = dict(routes or {})
        self.calls = []
        self.api_key = api_key
        self.config = Config(writable=False)

    def request(self, method, endpoint, params=None, json=None, timeout=None):
        self.calls.append((method, endpoint, params, json))
        hit = self.routes.get((method, endpoint))
        if hit is None:
            from prime_b200.platform.core import APIError

            raise APIError(f"HTTP 404: no route {method} {endpoint}", 404)
        if isinstance(hit, Exception):
            raise hit
        return hit(params=params, json=json) if callable(hit) else hit

    def get(self, endpoint, params=None, **kw):
        return self.request("GET", endpoint, params=params)

    def post(self, endpoint, json=None, params=None, **kw):
        return self.request("POST", endpoint, params=params, json=json)

    def patch(self, endpoint, json=None, **kw):
        return self.request("PATCH", endpoint, json=json)

    def put(self, endpoint, json=None, **kw):
        return self.request("PUT", endpoint, json=json)

    def delete(self, endpoint, params=None, json=None, **kw):
        return self.request("DELETE", endpoint, params=params, json=json)

    def called(self, method, endpoint):
        return [c for c in self.calls if c[0] == method and c[1] == endpoint]


@pytest.fixture
def fake_api(monkeypatch):
    """fake_api(routes, *modules) → FakeAPI installed as ``api()`` in each given command module."""

    def install(routes, *modules):
        fake = FakeAPI(routes)
        for m in modules:
            monkeypatch.setattr(m, "api", lambda require_auth=True, _f=fake: _f)
        return fake

    return install
