"""`[era5-svd]` / `[era5-download]` configuration contract of the reference
(src/dmd_era5/config_parser.py:51-290, config.ini:44-68, config_reader.py:16-62), kept verbatim in
meaning so that the DVC / era5_svd plumbing can call the B200 stage unchanged: same required fields,
same literal types, same derived names and paths, same error texts the reference's tests match on
(tests/test_03_era5_svd.py:104-150).  Table-driven restatement, not a copy.

Opt-in extension keys (absent = reference behaviour; the reference's parser ignores unknown keys):
    precision ("auto" | "native" | "tf32x3" | "tf32mix"), random_seed (int | None), area_weighting (bool), device (str),
    matrix_dtype ("float32" | "float64": the build kernel casts while it stacks; north_star "float cast"),
    n_gpus (int >= 1: row-shard the stage over that many GPUs of this node, stage.main).
Their values are validated here with the same error style as the reference's own fields.
"""
from __future__ import annotations

import ast
import configparser
import os
from datetime import datetime, timedelta

ERA5_PRESSURE_LEVEL_VARIABLES = {"temperature", "u_component_of_wind", "v_component_of_wind"}   # constants.py:5-9
ERA5_SINGLE_LEVEL_VARIABLES = {"2m_temperature", "10m_u_component_of_wind", "10m_v_component_of_wind"}   # :11-15
ERA5_PRESSURE_LEVELS = {50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000}        # :20-34

_COMMON = ["source_path", "start_datetime", "end_datetime", "delta_time", "variables", "levels"]
REQUIRED = {
    "era5-download": _COMMON,
    "era5-svd": ["source_path", "variables", "levels", "svd_type", "delay_embedding", "mean_center", "scale",
                 "start_datetime", "end_datetime", "delta_time", "n_components", "save_data_matrix"],
}
_DELTA_UNITS = {"h": lambda x: timedelta(hours=x), "d": lambda x: timedelta(days=x), "w": lambda x: timedelta(weeks=x),
                "m": lambda x: timedelta(days=x * 365 // 12), "y": lambda x: timedelta(days=x * 365)}
EXTENSION_KEYS = ("precision", "random_seed", "area_weighting", "device", "matrix_dtype", "n_gpus")
PRECISION_VALUES = ("auto", "native", "fp64", "fp32", "tf32x3", "tf32mix")


def _check_extension(key: str, val, logger=None):
    """Validate an opt-in key (a bad value used to surface as a bare KeyError deep inside the driver)."""
    if key == "precision" and val not in PRECISION_VALUES:
        _fail(f"precision {val} is not supported. Supported values are {', '.join(PRECISION_VALUES)}.", logger)
    if key == "matrix_dtype" and val not in (None, "float32", "float64"):
        _fail(f"matrix_dtype {val} is not supported. Supported values are float32, float64.", logger)
    if key == "random_seed" and not (val is None or (isinstance(val, int) and not isinstance(val, bool) and val >= 0)):
        _fail("random_seed must be None or a non-negative integer.", logger)
    if key == "area_weighting" and not isinstance(val, bool):
        _fail("area_weighting must be a boolean.", logger)
    if key == "n_gpus" and not (isinstance(val, int) and not isinstance(val, bool) and val >= 1):
        _fail("n_gpus must be an integer greater than 0.", logger)


def project_root() -> str:
    """Stand-in for pyprojroot.here(): $DMD_ERA5_ROOT or the current working directory."""
    return os.environ.get("DMD_ERA5_ROOT", os.getcwd())


def config_reader(section: str, config_path: str | None = None) -> dict:
    """INI section -> dict, every value through ast.literal_eval (config_reader.py:16-62), with the reference's error
    contract: a missing file reads as "no such section" (ConfigParser.read skips it), a missing section raises
    ``Exception("Section ... not found in the ... file")``, an unparsable value is logged / printed and the ORIGINAL
    exception of ast.literal_eval propagates."""
    import logging

    path = config_path or os.path.join(project_root(), "config.ini")
    log = logging.getLogger("ConfigReader")
    parser = configparser.ConfigParser()
    parser.read(path, encoding="utf-8-sig")
    out = {}
    if parser.has_section(section):
        for param_name, param_value in parser.items(section):
            try:
                out[param_name] = ast.literal_eval(param_value)
            except Exception as e:
                msg = f"""
                    Error while parsing {param_name} from {section} section
                    in the config file: {e}
                    """
                log.error(msg)
                print(msg)
                raise
    else:
        msg = f"Section {section} not found in the {path} file"
        log.error(msg)
        raise Exception(msg)
    return out


def _fail(msg: str, logger=None):
    if logger is not None:
        logger.error(msg)
    raise ValueError(msg)


def validate_time_parameters(parsed: dict) -> None:
    """config_parser.py:14-48."""
    start, end, delta = parsed["start_datetime"], parsed["end_datetime"], parsed["delta_time"]
    if end <= start:
        raise ValueError("End datetime must be after start datetime")
    if (end - start) < delta:
        raise ValueError(f"Time range must be at least as long as delta_time.\n        {end} - {start} < {delta}")
    if delta <= timedelta(0):
        raise ValueError("delta_time must be positive.")
    if start > datetime.now():
        raise ValueError("Start date cannot be in the future.")


def _typed(config, parsed, key, kind, what, rule, logger):
    val = config[key]
    ok = isinstance(val, bool) if kind is bool else (isinstance(val, int) and not isinstance(val, bool) and val >= 1)
    if not ok:
        _fail(f"\n            Invalid {what} in config: {val}.\n            {rule}\n            ", logger)
    parsed[key] = val


def config_parser(config: dict, section: str, logger=None) -> dict:
    if section not in REQUIRED:
        raise ValueError(f"Section {section} is not currently supported.")
    for field in REQUIRED[section]:
        if field not in config:
            _fail(f"Missing required field in config: {field}", logger)
    parsed: dict = {"source_path": config["source_path"]}
    try:
        parsed["start_datetime"] = datetime.fromisoformat(config["start_datetime"])
        parsed["end_datetime"] = datetime.fromisoformat(config["end_datetime"])
    except ValueError as e:
        _fail(f"Invalid datetime format in config: {e}", logger)
    try:
        unit, num = config["delta_time"][-1].lower(), int(config["delta_time"][:-1])
        if unit not in _DELTA_UNITS:
            raise ValueError(f"Unsupported delta_time format in config: {config['delta_time']}")
        parsed["delta_time"] = _DELTA_UNITS[unit](num)
    except ValueError as e:
        _fail(f"Error parsing delta_time from config: {e}", logger)
    validate_time_parameters(parsed)

    try:
        if config["variables"] == "all_pressure_level_vars":
            parsed["variables"] = list(ERA5_PRESSURE_LEVEL_VARIABLES)     # set order: reference quirk Q2
        elif config["variables"] == "all_single_level_vars":
            raise ValueError("Single level variables not currently supported.")
        else:
            parsed["variables"] = [v.strip() for v in config["variables"].split(",")]
            for var in parsed["variables"]:
                if var in ERA5_SINGLE_LEVEL_VARIABLES:
                    raise ValueError(f"Single level variables not currently supported: {var}")
                if var not in ERA5_PRESSURE_LEVEL_VARIABLES:
                    raise ValueError(f"Unsupported variable in config: {var}")
    except ValueError as e:
        _fail(f"Error parsing variables from config: {e}", logger)
    try:
        if config["levels"] == "all":
            parsed["levels"] = list(ERA5_PRESSURE_LEVELS)
        else:
            parsed["levels"] = [int(v) for v in config["levels"].split(",")]
            for level in parsed["levels"]:
                if level not in ERA5_PRESSURE_LEVELS:
                    raise ValueError(f"Unsupported level in config: {level}")
    except ValueError as e:
        _fail(f"Error parsing levels from config: {e}", logger)

    root = project_root()
    sub = "era5_download" if section == "era5-download" else "era5_svd"
    name = "{}_{}_{}.nc".format(parsed["start_datetime"].strftime("%Y-%m-%dT%H"),
                                parsed["end_datetime"].strftime("%Y-%m-%dT%H"), config["delta_time"])
    parsed["save_name"] = name
    parsed["save_path"] = os.path.join(root, "data", sub, name)
    parsed["era5_slice_path"] = os.path.join(root, "data", "era5_download", name)

    if section == "era5-svd":
        parsed["era5_svd_path"] = os.path.join(root, "data", "era5_svd", name)
        parsed["svd_type"] = config["svd_type"]
        supported = ["standard", "randomized"]
        if parsed["svd_type"] not in supported:
            _fail(f"\n            Invalid SVD type in config: {parsed['svd_type']}.\n            Supported types: {supported}.\n            ", logger)
        _typed(config, parsed, "delay_embedding", int, "delay embedding", "Delay embedding must be an integer greater than 0.", logger)
        _typed(config, parsed, "mean_center", bool, "mean centering", "Mean centering must be a boolean value.", logger)
        _typed(config, parsed, "scale", bool, "scaling", "Scaling must be a boolean value.", logger)
        _typed(config, parsed, "n_components", int, "number of components", "Number of components must be an integer greater than 0.", logger)
        _typed(config, parsed, "save_data_matrix", bool, "save_data_matrix", "save_data_matrix must be a boolean value.", logger)
        for key in EXTENSION_KEYS:
            if key in config:
                _check_extension(key, config[key], logger)
                parsed[key] = config[key]
    return parsed
