"""DVC plumbing around the SVD stage: versioned retrieval of slices / results and the YAML side-log that records
which configuration produced which DVC md5 (reference: src/dmd_era5/dvc_tools.py:11-253, called from
era5_svd.py:116-154, 190-227, 442-451).

`dvc` and `GitPython` are imported lazily: the log-file format and the version-matching rules (the part of the
plumbing with behaviour of its own) work - and are tested - without them; the calls that need a DVC repository raise
``DvcUnavailable`` when the packages are missing, which the stage reports exactly like a failed DVC retrieval in the
reference (warning, fall through to computing).  Same function names, arguments, error texts and return values.
"""
from __future__ import annotations

import os
import subprocess
from datetime import datetime

from .config_parser import project_root


class DvcUnavailable(FileNotFoundError):
    """dvc / GitPython are not importable in this environment (a FileNotFoundError: the reference's callers
    already treat that as "could not retrieve from DVC")."""


def _repos():
    try:
        from dvc.repo import Repo as DvcRepo
        from git import Repo as GitRepo
    except Exception as e:   # pragma: no cover - depends on the environment
        raise DvcUnavailable(f"DVC plumbing needs the 'dvc' and 'GitPython' packages ({e})") from e
    return DvcRepo, GitRepo


def dvc_available() -> bool:
    try:
        _repos()
        return True
    except DvcUnavailable:
        return False


def _load_yaml(path: str):
    import yaml

    with open(path) as f:
        return yaml.safe_load(f)


def add_config_to_dvc_log(dvc_file_path: str, data_path: str, data_attrs: dict, git_add: bool = False) -> None:
    """dvc_tools.py:11-47: append ``<md5 of the .dvc file's first out>:`` followed by one ``  key: value`` line per
    attribute to ``<data_path>.yaml`` (created when missing); optionally stage the log for commit."""
    md5_hash = _load_yaml(dvc_file_path)["outs"][0]["md5"]
    log_file = data_path + ".yaml"
    if not os.path.exists(log_file):
        with open(log_file, "w") as f:
            f.write("")
    with open(log_file, "a") as f:
        f.write(f"{md5_hash}:\n")
        for key, value in data_attrs.items():
            f.write(f"  {key}: {value}\n")
    if git_add:
        _, GitRepo = _repos()
        with GitRepo(project_root()) as repo:
            repo.index.add([log_file])


def add_data_to_dvc(data_path: str, data_attrs: dict) -> None:
    """dvc_tools.py:50-63: ``dvc add`` the file, then log its attributes under the new md5."""
    DvcRepo, _ = _repos()
    with DvcRepo(project_root()) as repo:
        repo.add(data_path)
    add_config_to_dvc_log(data_path + ".dvc", data_path, data_attrs, git_add=True)


def find_first_commit_with_md5_hash(md5_hash: str, dvc_file_path: str) -> str | None:
    """dvc_tools.py:66-92: oldest commit whose diff of the .dvc file mentions the md5 (``git log -S``)."""
    command = ["git", "log", "-S", md5_hash, "--reverse", "--oneline", "--", dvc_file_path]
    result = subprocess.run(command, stdout=subprocess.PIPE, text=True, check=False)
    lines = result.stdout.strip().splitlines()
    return lines[0].split()[0] if lines else None


def fetch_data_from_default_remote(repo, targets: list) -> tuple[bool, bool]:
    """dvc_tools.py:95-116: (a default remote exists, something was fetched)."""
    remotes = repo.config["remote"]
    if not remotes:
        return False, False
    return True, repo.fetch(targets=targets) > 0


def _paths(parsed_config: dict, data_type: str) -> tuple[str, str]:
    if data_type == "era5_slice":
        key, what = "era5_slice_path", "ERA5 slice"
    elif data_type == "era5_svd":
        key, what = "era5_svd_path", "ERA5 SVD data"
    else:
        raise ValueError("\n        Data type not supported.\n        Currently only 'era5_slice' and 'era5_svd' are supported.\n        ")
    try:
        base = parsed_config[key]
    except KeyError as e:
        raise KeyError(f"\n            The configuration dictionary does not contain the path to the {what}.\n"
                       "            Are you sure you are using the correct configuration or requesting\n"
                       "            the correct data type?\n            ") from e
    return base + ".yaml", base + ".dvc"


def select_logged_version(log_content: dict, parsed_config: dict, data_type: str) -> str | None:
    """The matching rules of dvc_tools.py:181-207 on the parsed side-log: the md5 of the MOST RECENT entry whose
    metadata satisfies the request.  Slices: superset match on variables and levels, equal source path, ordered by
    ``date_downloaded``.  Results: equal source path, variables, levels (ORDER-sensitive, quirk Q2), delay embedding,
    mean_center, scale and n_components - svd_type and save_data_matrix are not compared (quirk Q5) - ordered by
    ``date_processed``."""
    keep, date_keep = None, datetime(1970, 1, 1)
    for md5_hash, meta in (log_content or {}).items():
        if data_type == "era5_slice":
            ok = (sorted(parsed_config["variables"]) == sorted(set(meta["variables"]) & set(parsed_config["variables"]))
                  and sorted(parsed_config["levels"]) == sorted(set(meta["levels"]) & set(parsed_config["levels"]))
                  and parsed_config["source_path"] == meta["source_path"])
            stamp = "date_downloaded"
        else:
            ok = (parsed_config["source_path"] == meta["source_path"] and parsed_config["variables"] == meta["variables"]
                  and parsed_config["levels"] == meta["levels"]
                  and parsed_config["delay_embedding"] == meta["delay_embedding"]
                  and parsed_config["mean_center"] == bool(meta["mean_center"])
                  and parsed_config["scale"] == bool(meta["scale"])
                  and parsed_config["n_components"] == meta["n_components"])
            stamp = "date_processed"
        if ok and meta[stamp] > date_keep:
            keep, date_keep = md5_hash, meta[stamp]
    return keep


_NO_RETRIEVE = ("\n            Found a matching version of the data in the log file,\n"
                "            but could not retrieve it from DVC.\n            ")


# the same sentence one block deeper in the reference's source (dvc_tools.py:243-247): its text carries that indentation
_NO_RETRIEVE_AFTER_FETCH = ("\n                    Found a matching version of the data in the log file,\n"
                            "                    but could not retrieve it from DVC.\n                    ")


def retrieve_data_from_dvc(parsed_config: dict, data_type: str = "era5_slice") -> None:
    """dvc_tools.py:119-253: check out the most recent logged version of the data that matches the configuration.
    FileNotFoundError when the .dvc or log file is missing, ValueError when nothing matches or the matching version
    cannot be brought back (not in the local cache and no default remote / nothing fetched)."""
    log_file_path, dvc_file_path = _paths(parsed_config, data_type)
    if not os.path.exists(log_file_path) or not os.path.exists(dvc_file_path):
        raise FileNotFoundError("DVC file or log file does not exist.")
    md5 = select_logged_version(_load_yaml(log_file_path), parsed_config, data_type)
    if not md5:
        raise ValueError("No matching version of the data found in DVC.")
    commit_hash = find_first_commit_with_md5_hash(md5, dvc_file_path)
    if commit_hash is None:
        raise ValueError(_NO_RETRIEVE)
    DvcRepo, GitRepo = _repos()
    root = project_root()
    with GitRepo(root) as repo:
        repo.git.checkout(commit_hash, dvc_file_path)
    with DvcRepo(root) as repo:
        if os.path.exists(os.path.join(root, ".dvc/cache/files/md5", md5[:2], md5[2:])):
            print("Checked out files:", repo.checkout(targets=[dvc_file_path]))
            return
        print("\n                    Data not found in local DVC cache.\n"
              "                    Attempting to fetch from default remote.\n                    ")
        remote_exists, data_fetched = fetch_data_from_default_remote(repo, targets=[dvc_file_path])
        if data_fetched:
            print("Data successfully fetched.")
            print("Checked out files:", repo.checkout(targets=[dvc_file_path]))
        if not remote_exists or not data_fetched:
            print("Could not fetch data from default remote DVC repository.")
            raise ValueError(_NO_RETRIEVE_AFTER_FETCH)
