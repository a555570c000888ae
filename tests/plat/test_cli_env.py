"""``prime env``: archive contents + content hash, version bumps, tar safety, install-command construction, and the
push / pull / install / list / actions / secrets flows against a fake API
(reference tests: packages/prime/tests/test_env_*.py, test_content_hash.py, test_install_*.py)."""

import io
import json
import tarfile
import zipfile
from pathlib import Path

import httpx
import pytest
from typer.testing import CliRunner

from prime_b200.platform.commands import env as env_mod
from prime_b200.platform.commands import env_packaging as pk
from prime_b200.platform.commands import env_secrets as sec_mod
from prime_b200.platform.core import APIError
from prime_b200.platform.main import app

runner = CliRunner()


def make_env(root: Path, name="my-env", version="0.1.0") -> Path:
    root.mkdir(parents=True, exist_ok=True)
    (root / "pyproject.toml").write_text(f'[project]\nname = "{name}"\nversion = "{version}"  # keep\ndescription = "d"\n\n[tool.x]\nversion = "9"\n')
    (root / "README.md").write_text("# hi\n")
    (root / "my_env.py").write_text("def load_environment():\n    return 1\n")
    (root / "pkg").mkdir()
    (root / "pkg" / "__init__.py").write_text("")
    (root / "pkg" / "data.json").write_text("{}")
    return root


# ----------------------------------------------------------------------------------------------- packaging helpers
def test_archive_selection_and_hash(tmp_path):
    e = make_env(tmp_path / "e")
    (e / ".hidden").write_text("x")
    (e / "notes.txt").write_text("top-level non-py is not part of the env")
    for junk in ("dist", "build", "outputs", "__pycache__", "x.egg-info", ".git"):
        (e / junk).mkdir()
        (e / junk / "f.py").write_text("junk")
    (e / "pkg" / "__pycache__").mkdir()
    (e / "pkg" / "__pycache__" / "a.pyc").write_text("junk")
    (e / "pkg" / "link.py").symlink_to(e / "my_env.py")
    rel = [p.relative_to(e).as_posix() for p in pk.collect_archive_files(e)]
    assert rel == ["README.md", "my_env.py", "pkg/__init__.py", "pkg/data.json", "pyproject.toml"]
    h1 = pk.compute_content_hash(e)
    (e / "dist" / "g.py").write_text("more junk")
    assert pk.compute_content_hash(e) == h1  # ignored dirs do not affect the hash
    (e / "pkg" / "data.json").write_text("{ }")
    assert pk.compute_content_hash(e) != h1
    # renaming a file changes the hash even when bytes are identical
    h2 = pk.compute_content_hash(e)
    (e / "my_env.py").rename(e / "my_env2.py")
    assert pk.compute_content_hash(e) != h2


def test_gitignore_is_honoured(tmp_path):
    e = make_env(tmp_path / "e")
    (e / "pkg" / "big.bin").write_text("weights")
    (e / "cache").mkdir()
    (e / "cache" / "x.py").write_text("1")
    (e / ".gitignore").write_text("*.bin\ncache/\n")
    rel = [p.relative_to(e).as_posix() for p in pk.collect_archive_files(e)]
    assert "pkg/big.bin" not in rel and "cache/x.py" not in rel and "pkg/data.json" in rel


def test_version_bumps_and_pyproject_rewrite(tmp_path):
    assert pk.bump_version("1.2.3") == "1.2.4" and pk.bump_version("1.2.3rc1") == "1.2.4"
    assert pk.bump_rc_version("1.2.3") == "1.2.3rc1" and pk.bump_rc_version("1.2.3rc1") == "1.2.3rc2"
    assert pk.bump_post_version("1.2.3") == "1.2.3.post1" and pk.bump_post_version("1.2.3.post4") == "1.2.3.post5"
    with pytest.raises(ValueError):
        pk.bump_version("banana")
    e = make_env(tmp_path / "e")
    pk.update_pyproject_version(e / "pyproject.toml", "0.1.1")
    text = (e / "pyproject.toml").read_text()
    assert 'version = "0.1.1"  # keep' in text and '[tool.x]\nversion = "9"' in text


def test_validate_env_id_and_slug():
    assert pk.validate_env_id("a/b") == ("a/b", "latest") and pk.validate_env_id("a/b@1.0") == ("a/b", "1.0")
    for bad in ("a", "a/b/c", "/b", "a/b@"):
        with pytest.raises(ValueError):
            pk.validate_env_id(bad)
    assert pk.parse_environment_slug("o/n@2") == ("o", "n")
    assert pk.normalize_package_name("My_Env.x") == "my-env-x"


def _tar(members):
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as t:
        for info, data in members:
            t.addfile(info, io.BytesIO(data) if data is not None else None)
    buf.seek(0)
    return tarfile.open(fileobj=buf, mode="r:gz")


def _file(name, data=b"x"):
    ti = tarfile.TarInfo(name)
    ti.size = len(data)
    return ti, data


def test_safe_tar_extract_refuses_escapes(tmp_path):
    ok = _tar([_file("a/b.py"), _file("c.txt")])
    pk.safe_tar_extract(ok, tmp_path / "ok")
    assert (tmp_path / "ok" / "a" / "b.py").read_bytes() == b"x"
    sym = tarfile.TarInfo("evil")
    sym.type, sym.linkname = tarfile.SYMTYPE, "/tmp"
    hard = tarfile.TarInfo("h")
    hard.type, hard.linkname = tarfile.LNKTYPE, "c.txt"
    for members in ([(sym, None)], [(hard, None)], [_file("/abs.txt")], [_file("../up.txt")], [_file("a/../../up.txt")]):
        dest = tmp_path / "bad"
        with pytest.raises(ValueError):
            pk.safe_tar_extract(_tar(members), dest)
        assert not dest.exists() or not any(dest.iterdir())  # nothing extracted before validation finished
    for comp in ("..", "a/b", "a\\b", "", "x\x00"):
        with pytest.raises(ValueError):
            pk.validate_path_component(comp, "owner")


def test_install_command_construction(monkeypatch):
    monkeypatch.setattr(pk.sys, "executable", "/venv/bin/python")
    idx = "https://hub/simple"
    assert pk.build_install_command("My_Env", "1.0", idx, None) == ["uv", "pip", "install", "--python", "/venv/bin/python", "-P", "my-env", "my-env==1.0", "--extra-index-url", idx]
    assert pk.build_install_command("e", "latest", idx, None, tool="pip", no_upgrade=True) == ["pip", "install", "e", "--extra-index-url", idx]
    assert pk.build_install_command("e", "latest", None, "https://h/e.whl", url_dependencies=["dep @ https://x/y.whl"])[-2:] == ["https://h/e.whl", "dep @ https://x/y.whl"]
    assert pk.build_install_command("e", "latest", None, "not-a-url") is None and pk.build_install_command("e", "1", None, None) is None
    with pytest.raises(ValueError):
        pk.get_install_command("conda", "https://h/e.whl", "e")


def test_is_environment_installed_checks_version():
    class R:
        def __init__(self, rc, out=""):
            self.returncode, self.stdout = rc, out

    seen = []
    run = lambda cmd, **kw: (seen.append(cmd), R(0, "Name: my-env\nVersion: 1.2.0\n"))[1]  # noqa: E731
    assert pk.is_environment_installed("my_env", None, runner=run) and seen[0][-1] == "my-env"
    assert pk.is_environment_installed("my_env", "1.2.0", runner=run) and not pk.is_environment_installed("my_env", "1.3.0", runner=run)
    assert pk.is_environment_installed("my_env", "latest", runner=run)
    assert not pk.is_environment_installed("x", None, runner=lambda cmd, **kw: R(1))


def test_requires_dist_from_wheel(tmp_path):
    w = tmp_path / "e-1-py3-none-any.whl"
    with zipfile.ZipFile(w, "w") as z:
        z.writestr("e-1.dist-info/METADATA", "Name: e\nRequires-Dist: verifiers>=0.1\nRequires-Dist: dep @ https://x/y.whl\n")
    assert pk.extract_requires_dist_from_wheel(w) == ["verifiers>=0.1", "dep @ https://x/y.whl"]
    assert pk.extract_requires_dist_from_wheel(tmp_path / "missing.whl") == []


def test_new_log_lines_overlap():
    assert env_mod.new_log_lines("", "a\nb") == ["a", "b"]
    assert env_mod.new_log_lines("a\nb\nc", "b\nc\nd\ne") == ["d", "e"]
    assert env_mod.new_log_lines("a\nb", "x\ny") == ["x", "y"]


# ----------------------------------------------------------------------------------------------- list / status / info
ENVS = {"data": [{"owner": {"name": "acme"}, "name": "math", "description": "sums", "visibility": "PUBLIC", "latest_version": "0.2.0",
                  "stars": 3, "updated_at": "2026-01-02T03:04:05Z", "latest_ci_status": "SUCCESS", "tags": ["t"]}], "total_count": 41}  # fmt: skip


def test_env_list_params_and_json(fake_api):
    api = fake_api({("GET", "/environmentshub/"): ENVS}, env_mod)
    r = runner.invoke(app, ["env", "list", "--output", "json", "--search", "ma", "-t", "t", "--page", "2", "-n", "20", "--show-actions", "--mine"])
    assert r.exit_code == 0, r.output
    out = json.loads(r.output)
    assert out["total"] == 41 and out["page"] == 2 and out["environments"][0] == {
        "environment": "acme/math", "description": "sums", "visibility": "PUBLIC", "version": "0.2.0", "stars": 3,
        "updated_at": "2026-01-02T03:04:05Z", "action_status": "SUCCESS", "tags": ["t"]}  # fmt: skip
    p = api.calls[0][2]
    assert p["offset"] == 20 and p["limit"] == 20 and p["search"] == "ma" and p["tags"] == ["t"] and p["include_ci_status"] and p["mine_only"]
    assert "starred_only" not in p and p["sort_by"] == "created_at"
    r = runner.invoke(app, ["env", "list"])
    assert "acme/math" in r.output and "Use --page 2" in r.output and "2026-01-02" in r.output
    assert runner.invoke(app, ["env", "list", "--sort", "bogus"]).exit_code == 1
    assert runner.invoke(app, ["env", "list", "--page", "0"]).exit_code == 1


def test_env_status_and_info(fake_api):
    fake_api({("GET", "/environmentshub/acme/math/status"): {"data": {"name": "math", "visibility": "PUBLIC", "latest_version": {
                  "semantic_version": "0.2.0", "content_hash": "abcdef0123456789", "created_at": "2026-01-02T03:04:05Z"},
                  "action": {"status": "FAILED", "job_id": "j1"}}},
              ("GET", "/environmentshub/acme/math/@0.2.0"): {"data": {"simple_index_url": "https://hub/simple", "metadata": {"description": "sums"}}},
              ("GET", "/environmentshub/acme/priv/@latest"): {"data": {"visibility": "PRIVATE"}}}, env_mod)  # fmt: skip
    r = runner.invoke(app, ["env", "status", "acme/math"])
    assert r.exit_code == 0 and "0.2.0" in r.output and "abcdef012345" in r.output and "FAILED" in r.output and "j1" in r.output
    r = runner.invoke(app, ["env", "info", "acme/math@0.2.0"])
    assert r.exit_code == 0 and "uv pip install math==0.2.0 --extra-index-url https://hub/simple" in r.output
    r = runner.invoke(app, ["env", "info", "acme/priv"])
    assert "prime env pull acme/priv@latest" in r.output
    assert runner.invoke(app, ["env", "info", "nope"]).exit_code == 1


# ----------------------------------------------------------------------------------------------- push
def _push_routes(created=True, finalize_ok=True):
    return {("POST", "/environmentshub/resolve"): {"data": {"id": "E1", "owner": {"name": "acme"}, "created": created}},
            ("POST", "/environmentshub/E1/wheels"): {"data": {"wheel_id": "W1", "upload_url": "https://s3/wheel"}},
            ("POST", "/environmentshub/E1/wheels/W1/finalize"): {"data": {}},
            ("POST", "/environmentshub/E1/versions"): {"data": {"version_id": "V1", "upload_url": "https://s3/src"}},
            ("POST", "/environmentshub/E1/versions/V1/finalize"): {"data": {"success": finalize_ok, "message": "nope"}}}  # fmt: skip


@pytest.fixture
def fake_build(monkeypatch):
    def build(env_path):
        d = env_path / "dist"
        d.mkdir(exist_ok=True)
        w = d / "my_env-0.1.1-py3-none-any.whl"
        with zipfile.ZipFile(w, "w") as z:
            z.writestr("my_env-0.1.1.dist-info/METADATA", "Requires-Dist: verifiers\n")
        return w

    monkeypatch.setattr(pk, "build_wheel", build)
    puts = []
    monkeypatch.setattr(env_mod.httpx, "put", lambda url, content=None, **kw: (puts.append((url, len(content))), httpx.Response(200, request=httpx.Request("PUT", url)))[1])
    return puts


def test_push_full_flow(tmp_path, fake_api, fake_build, monkeypatch):
    e = make_env(tmp_path / "environments" / "my_env")
    (e / ".env-metadata.json").write_text(json.dumps({"owner": "old", "name": "my-env", "extra": 1}))  # legacy location
    api = fake_api(_push_routes(), env_mod)
    monkeypatch.chdir(tmp_path)
    r = runner.invoke(app, ["env", "push", "my-env", "--auto-bump", "--visibility", "PRIVATE", "--team", "tm"])
    assert r.exit_code == 0, r.output
    assert "0.1.0 → 0.1.1" in r.output and "Successfully pushed acme/my-env" in r.output and "Upstream set to acme/my-env" in r.output
    assert api.calls[0][3] == {"name": "my-env", "visibility": "PRIVATE", "team_slug": "tm"}
    wheel_req = api.called("POST", "/environmentshub/E1/wheels")[0][3]
    assert wheel_req["semantic_version"] == "0.1.1" and wheel_req["content_hash"] == pk.compute_content_hash(e)
    assert wheel_req["metadata"]["requires_dist"] == ["verifiers"] and len(wheel_req["sha256"]) == 64
    src_req = api.called("POST", "/environmentshub/E1/versions")[0][3]
    assert src_req["filename"] == f"my-env-0.1.1-{wheel_req['content_hash'][:8]}.tar.gz" and src_req["metadata"]["original_filename"] == "my-env-0.1.1.tar.gz"
    assert [u for u, _ in fake_build] == ["https://s3/wheel", "https://s3/src"]
    md = json.loads((e / ".prime" / ".env-metadata.json").read_text())
    assert md["owner"] == "acme" and md["environment_id"] == "E1" and md["extra"] == 1 and not (e / ".env-metadata.json").exists()


def test_push_failures(tmp_path, fake_api, fake_build, monkeypatch):
    monkeypatch.chdir(tmp_path)
    assert "pyproject.toml not found" in runner.invoke(app, ["env", "push"]).output
    e = make_env(tmp_path / "e")
    r = runner.invoke(app, ["env", "push", "-p", str(e), "--rc", "--post"])
    assert r.exit_code == 1 and "mutually exclusive" in r.output
    routes = _push_routes()
    routes[("POST", "/environmentshub/E1/wheels")] = APIError("Content hash abc already exists")
    fake_api(routes, env_mod)
    r = runner.invoke(app, ["env", "push", "-p", str(e)])
    assert r.exit_code == 1 and "--auto-bump" in r.output
    fake_api(_push_routes(finalize_ok=False), env_mod)
    r = runner.invoke(app, ["env", "push", "-p", str(e)])
    assert r.exit_code == 1 and "Error finalizing: nope" in r.output
    bare = tmp_path / "bare"
    bare.mkdir()
    (bare / "pyproject.toml").write_text('[project]\nname="x"\nversion="1.0.0"\n')
    assert "No environment Python file found" in runner.invoke(app, ["env", "push", "-p", str(bare)]).output


def test_push_prompts_for_username_once(tmp_path, fake_api, fake_build):
    e = make_env(tmp_path / "e")
    state = {"n": 0}

    def resolve(params=None, json=None):
        state["n"] += 1
        if state["n"] == 1:
            raise APIError("Your profile is missing a username")
        return {"data": {"id": "E1", "owner": {"name": "neo"}, "created": True}}

    def slug(params=None, json=None):
        if json["slug"] == "taken":
            raise APIError("HTTP 409: already taken")
        return {}

    routes = _push_routes()
    routes[("POST", "/environmentshub/resolve")] = resolve
    routes[("PATCH", "/user/slug")] = slug
    api = fake_api(routes, env_mod)
    r = runner.invoke(app, ["env", "push", "-p", str(e)], input="Bad Name!\ntaken\nneo\n")
    assert r.exit_code == 0, r.output
    assert "Invalid username" in r.output and "already taken" in r.output and "Username set to neo" in r.output
    assert [c[3]["slug"] for c in api.called("PATCH", "/user/slug")] == ["taken", "neo"]


# ----------------------------------------------------------------------------------------------- pull / install
def _source_tgz(tmp_path, version="0.3.0"):
    src = make_env(tmp_path / "src", version=version)
    p = tmp_path / "src.tar.gz"
    pk.build_source_archive(src, p)
    return p.read_bytes()


def test_pull_extracts_and_writes_metadata(tmp_path, fake_api, monkeypatch):
    blob = _source_tgz(tmp_path)
    seen = {}

    def dl(url, dest, key):
        seen.update(url=url, key=key)
        Path(dest).write_bytes(blob)

    monkeypatch.setattr(env_mod, "_download", dl)
    fake_api({("GET", "/environmentshub/acme/math/@0.3.0"): {"data": {"id": "E9", "package_url": "https://s3/pkg"}},
              ("GET", "/environmentshub/acme/none/@latest"): {"data": {}}}, env_mod)  # fmt: skip
    monkeypatch.chdir(tmp_path)
    (tmp_path / "math").mkdir()
    r = runner.invoke(app, ["env", "pull", "acme/math@0.3.0"])
    assert r.exit_code == 0, r.output
    dest = tmp_path / "math-1"  # ./math existed → next free suffix
    assert (dest / "pkg" / "data.json").exists() and seen == {"url": "https://s3/pkg", "key": "k-test"}
    assert json.loads((dest / ".prime" / ".env-metadata.json").read_text()) == {"environment_id": "E9", "owner": "acme", "name": "math"}
    assert "No downloadable package" in runner.invoke(app, ["env", "pull", "acme/none"]).output
    assert runner.invoke(app, ["env", "pull", "bad"]).exit_code == 1


def test_private_env_is_built_once_into_wheel_cache(tmp_path, fake_api, monkeypatch, isolated_home):
    blob = _source_tgz(tmp_path, version="0.3.0")
    downloads, builds = [], []
    monkeypatch.setattr(env_mod, "_download", lambda url, dest, key: (downloads.append(url), Path(dest).write_bytes(blob)))

    def build(slot):
        builds.append(slot)
        (slot / "dist").mkdir()
        w = slot / "dist" / "my_env-0.3.0-py3-none-any.whl"
        w.write_bytes(b"whl")
        return w

    monkeypatch.setattr(pk, "build_wheel", build)
    api = fake_api({}, env_mod)
    details = {"id": "E2", "package_url": "https://s3/p", "visibility": "PRIVATE"}
    wheel, ver = env_mod.pull_and_build_private_env(api, "acme", "my-env", "latest", details)
    slot = isolated_home / ".prime" / "wheel_cache" / "acme" / "my-env" / "0.3.0"
    assert ver == "0.3.0" and wheel.parent == slot / "dist" and (slot / "pyproject.toml").exists()
    assert json.loads((slot / ".prime" / ".env-metadata.json").read_text())["version"] == "0.3.0"
    # "latest" must re-download to learn the version, but the cached wheel is reused (no second build)
    env_mod.pull_and_build_private_env(api, "acme", "my-env", "latest", details)
    assert len(downloads) == 2 and len(builds) == 1
    # a pinned version hits the cache without any network
    assert env_mod.pull_and_build_private_env(api, "acme", "my-env", "0.3.0", details) == (wheel, "0.3.0") and len(downloads) == 2
    with pytest.raises(ValueError):
        env_mod.pull_and_build_private_env(api, "..", "my-env", "latest", details)


def test_install_resolution_paths(tmp_path, fake_api, monkeypatch):
    ran = []
    monkeypatch.setattr(env_mod, "execute_install_command", lambda cmd, env_id, ver, tool: ran.append((cmd, env_id, ver, tool)))
    monkeypatch.setattr(env_mod.shutil, "which", lambda t: "/usr/bin/" + t)
    monkeypatch.setattr(pk.sys, "executable", "/venv/bin/python")
    fake_api({("GET", "/environmentshub/acme/idx/@1.0"): {"data": {"simple_index_url": "https://hub/simple", "url_dependencies": ["d @ https://x/d.whl"]}},
              ("GET", "/environmentshub/acme/whl/@latest"): {"data": {"wheel_url": "https://hub/w.whl"}},
              ("GET", "/environmentshub/acme/none/@latest"): {"data": {"visibility": "PUBLIC"}}}, env_mod)  # fmt: skip
    local = tmp_path / "environments" / "loc_env"
    local.mkdir(parents=True)
    r = runner.invoke(app, ["env", "install", "acme/idx@1.0", "acme/whl", "acme/none", "acme/missing", "bad/format/x", "loc-env", "ghost",
                            "acme/whl", "--with", "pip", "-p", str(tmp_path / "environments")])  # fmt: skip
    assert r.exit_code == 0, r.output
    assert [c[1:] for c in ran] == [("acme/idx", "1.0", "pip"), ("acme/whl", "latest", "pip"), ("loc-env", "local", "pip")]  # de-duplicated, in order
    assert ran[0][0] == ["pip", "install", "--upgrade", "idx==1.0", "d @ https://x/d.whl", "--extra-index-url", "https://hub/simple"]
    assert ran[1][0] == ["pip", "install", "--upgrade", "https://hub/w.whl"] and ran[2][0] == ["pip", "install", "-e", str(local)]
    assert "Skipping acme/none" in r.output and "Failed to resolve acme/missing" in r.output and "Local environment not found" in r.output
    assert "Installed 3 environments" in r.output
    assert runner.invoke(app, ["env", "install", "acme/none"]).exit_code == 1
    assert runner.invoke(app, ["env", "install", "acme/whl", "--with", "conda"]).exit_code == 1
    # library helper used by `prime eval run`
    ran.clear()
    assert env_mod.install_single_environment("acme/whl") and ran[0][0][:3] == ["uv", "pip", "install"]
    assert not env_mod.install_single_environment("acme/none") and not env_mod.install_single_environment("nope")


def test_uninstall_strips_owner(monkeypatch):
    seen = []

    class R:
        returncode, stdout, stderr = 0, "Uninstalled 1 package", ""

    monkeypatch.setattr(env_mod.shutil, "which", lambda t: "/usr/bin/" + t)
    monkeypatch.setattr(env_mod.subprocess, "run", lambda cmd, **kw: (seen.append(cmd), R())[1])
    r = runner.invoke(app, ["env", "uninstall", "acme/My_Env", "--with", "pip"])
    assert r.exit_code == 0 and seen == [["pip", "uninstall", "-y", "my-env"]] and "Successfully uninstalled my-env" in r.output


# ----------------------------------------------------------------------------------------------- versions / delete / actions
def test_versions_and_delete(fake_api):
    api = fake_api({("GET", "/environmentshub/acme/math/versions"): {"data": {"versions": [
                        {"version": "0.2.0", "created_at": "2026-01-02T03:04:05Z", "sha256": "a" * 64, "size": 2},
                        {"version": "0.1.0", "created_at": "2025-01-02", "sha256": "b" * 64, "size": 1}]}},
                    ("DELETE", "/environmentshub/acme/math/@aaaaaaaa"): {}, ("DELETE", "/environmentshub/acme/math"): {}}, env_mod)  # fmt: skip
    r = runner.invoke(app, ["env", "version", "list", "acme/math"])
    assert r.exit_code == 0 and "aaaaaaaa" in r.output and "a" * 9 not in r.output and "2 artifacts" in r.output and "1 artifact" in r.output
    assert "prime env install acme/math@0.2.0" in r.output
    assert "a" * 64 in runner.invoke(app, ["env", "version", "list", "acme/math", "--full-hashes"]).output
    assert runner.invoke(app, ["env", "version", "delete", "acme/math", "abc"]).exit_code == 1  # hash too short
    r = runner.invoke(app, ["env", "version", "delete", "acme/math", "aaaaaaaa"], input="n\n")
    assert "cancelled" in r.output and not api.called("DELETE", "/environmentshub/acme/math/@aaaaaaaa")
    assert runner.invoke(app, ["env", "version", "delete", "acme/math", "aaaaaaaa", "-f"]).exit_code == 0
    r = runner.invoke(app, ["env", "version", "delete", "acme/math", "cccccccc", "-f"])
    assert r.exit_code == 1 and "not found in environment" in r.output
    assert runner.invoke(app, ["env", "delete", "acme/math", "--force"]).exit_code == 0 and api.called("DELETE", "/environmentshub/acme/math")


def test_actions(fake_api):
    api = fake_api({("GET", "/environmentshub/acme/math/actions"): {"data": {"total": 30, "actions": [
                        {"id": "A1", "job_type": "integration", "status": "RUNNING", "version": {"content_hash": "deadbeefcafe"},
                         "trigger": "push", "created_at": "2026-01-02T03:04:05Z"}]}},
                    ("GET", "/environmentshub/acme/math/actions/A1/logs"): {"data": {"logs": "\x1b[31mred\x1b[0m line"}},
                    ("POST", "/environmentshub/acme/math/actions/retry"): lambda params=None, json=None: {"data": {"success": bool(json), "job_id": "J2", "message": "no action"}}},
                   env_mod)  # fmt: skip
    r = runner.invoke(app, ["env", "action", "list", "acme/math", "-v", "V1"])
    assert r.exit_code == 0 and "integration" in r.output and "deadbeef" in r.output and "Use --page 2" in r.output
    assert api.calls[0][2] == {"limit": 20, "offset": 0, "version_id": "V1"}
    r = runner.invoke(app, ["env", "action", "logs", "acme/math", "A1", "-n", "5"])
    assert r.output.strip() == "red line" and api.calls[-1][2] == {"tail_lines": 5}
    assert "J2" in runner.invoke(app, ["env", "action", "retry", "acme/math", "A1"]).output
    r = runner.invoke(app, ["env", "action", "retry", "acme/math"])
    assert r.exit_code == 1 and "no action" in r.output


# ----------------------------------------------------------------------------------------------- secrets / variables
def _secret_routes():
    return {("GET", "/environmentshub/acme/math/@latest"): {"data": {"id": "E1"}},
            ("GET", "/environmentshub/E1/secrets"): {"data": [{"id": "S1", "name": "API_KEY", "source": "environment", "createdAt": "2026-01-02T03:04:05Z"}]},
            ("POST", "/environmentshub/E1/secrets"): lambda params=None, json=None: {"data": {"id": "S2", **json}},
            ("PATCH", "/environmentshub/E1/secrets/S1"): lambda params=None, json=None: {"data": {"id": "S1", "name": json.get("name", "API_KEY")}},
            ("DELETE", "/environmentshub/E1/secrets/S1"): {}, ("POST", "/environmentshub/E1/secrets/link/G1"): {"data": {"secretName": "GLOBAL"}},
            ("DELETE", "/environmentshub/E1/secrets/link/G1"): {},
            ("GET", "/environmentshub/E1/variables"): {"data": [{"id": "V1", "name": "MODE", "value": "x" * 40}]},
            ("POST", "/environmentshub/E1/variables"): lambda params=None, json=None: {"data": {"id": "V2", **json}},
            ("PATCH", "/environmentshub/E1/variables/V1"): lambda params=None, json=None: {"data": {"id": "V1", "name": "MODE", **json}},
            ("DELETE", "/environmentshub/E1/variables/V1"): {}}  # fmt: skip


def test_env_secrets(tmp_path, fake_api, monkeypatch):
    api = fake_api(_secret_routes(), env_mod, sec_mod)
    r = runner.invoke(app, ["env", "secret", "list", "acme/math"])
    assert r.exit_code == 0 and "API_KEY" in r.output and "environment" in r.output
    assert json.loads(runner.invoke(app, ["env", "secret", "list", "acme/math", "-o", "json"]).output)["secrets"][0]["id"] == "S1"
    r = runner.invoke(app, ["env", "secret", "create", "acme/math", "-n", "TOKEN", "-v", "s3cr3t", "-d", "desc"])
    assert r.exit_code == 0 and "S2" in r.output and "s3cr3t" not in r.output
    assert api.called("POST", "/environmentshub/E1/secrets")[0][3] == {"name": "TOKEN", "value": "s3cr3t", "description": "desc"}
    assert runner.invoke(app, ["env", "secret", "create", "acme/math", "-n", "lower", "-v", "x"]).exit_code == 1
    # interactive: pick #1, then supply a hidden new value
    r = runner.invoke(app, ["env", "secret", "update", "acme/math"], input="1\nnewval\n")
    assert r.exit_code == 0 and api.called("PATCH", "/environmentshub/E1/secrets/S1")[-1][3] == {"value": "newval"}
    r = runner.invoke(app, ["env", "secret", "delete", "acme/math", "--id", "S1"], input="y\n")
    assert r.exit_code == 0 and "Deleted secret 'API_KEY'" in r.output
    assert "GLOBAL" in runner.invoke(app, ["env", "secret", "link", "G1", "acme/math"]).output
    assert runner.invoke(app, ["env", "secret", "unlink", "G1", "acme/math", "-y"]).exit_code == 0
    # slug auto-detected from ./.prime/.env-metadata.json
    (tmp_path / ".prime").mkdir()
    (tmp_path / ".prime" / ".env-metadata.json").write_text(json.dumps({"owner": "acme", "name": "math"}))
    monkeypatch.chdir(tmp_path)
    r = runner.invoke(app, ["env", "secret", "list"])
    assert r.exit_code == 0 and "Using environment: acme/math" in r.output
    monkeypatch.chdir(tmp_path.parent)
    assert runner.invoke(app, ["env", "secret", "list"]).exit_code == 1


def test_env_variables(fake_api):
    api = fake_api(_secret_routes(), env_mod, sec_mod)
    r = runner.invoke(app, ["env", "var", "list", "acme/math"])
    assert "MODE" in r.output and "x" * 27 + "..." in r.output and "x" * 28 not in r.output
    assert runner.invoke(app, ["env", "var", "create", "acme/math", "-n", "LEVEL", "-v", "3"]).exit_code == 0
    assert runner.invoke(app, ["env", "var", "update", "V1", "acme/math"]).exit_code == 1  # nothing to change
    r = runner.invoke(app, ["env", "var", "update", "V1", "acme/math", "-v", "y", "-o", "json"])
    assert json.loads(r.output) == {"id": "V1", "name": "MODE", "value": "y"}
    r = runner.invoke(app, ["env", "var", "delete", "V1", "acme/math"], input="n\n")
    assert not api.called("DELETE", "/environmentshub/E1/variables/V1")
    assert runner.invoke(app, ["env", "var", "delete", "V1", "acme/math", "-y"]).exit_code == 0
