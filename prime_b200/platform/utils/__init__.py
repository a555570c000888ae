from .config import BaseConfig, load_toml  # noqa: F401
from .display import build_table, output_data_as_json, validate_output_format  # noqa: F401
from .env_metadata import find_environment_metadata, get_environment_metadata  # noqa: F401
from .env_vars import EnvParseError, collect_env_vars, parse_env_arg, parse_env_file  # noqa: F401
from .formatters import (  # noqa: F401
    format_ip_display,
    format_price,
    format_resources,
    format_size,
    obfuscate_env_vars,
    obfuscate_secret,
    strip_ansi,
)
from .json_help import json_output_help, list_json_help  # noqa: F401
from .plain import HELP_NOTE, DefaultCommandGroup, PlainTyper, get_console, is_plain_mode  # noqa: F401
from .prompt import confirm_or_skip, prompt_for_value, require_selection, select_item_interactive, validate_env_var_name  # noqa: F401
from .time_utils import format_time_ago, human_age, iso_timestamp, now_utc, sort_by_created, to_utc  # noqa: F401
