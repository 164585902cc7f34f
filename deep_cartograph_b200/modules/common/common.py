"""Config + archive helpers with the semantics of the reference's ``modules/common/common.py``
(validate_configuration :195-232, merge_configurations :234-259, zip_files :72, unzip_files :126)."""
from __future__ import annotations

import logging
import os
import sys
import zipfile
from typing import Any, Dict, Optional, Type

import yaml
from pydantic import BaseModel, ValidationError

logger = logging.getLogger(__name__)


def validate_configuration(configuration: Dict[str, Any], schema: Type[BaseModel],
                           output_folder: Optional[str]) -> Dict[str, Any]:
    """Validate with the pydantic schema, dump ``configuration.yml`` next to the outputs."""
    try:
        validated = schema(**configuration).model_dump()
    except ValidationError as e:
        logger.error(f"Configuration file is not valid: {e}")
        sys.exit(1)
    if output_folder is not None:
        with open(os.path.join(output_folder, "configuration.yml"), "w") as fh:
            yaml.dump(validated, fh)
    return validated


def merge_configurations(common_config: Dict, specific_config: Optional[Dict]) -> Dict:
    """Recursive merge; CV-specific values win, common keys are preserved."""
    merged = dict(common_config)
    for key, value in (specific_config or {}).items():
        if isinstance(merged.get(key), dict) and isinstance(value, dict):
            merged[key] = merge_configurations(merged[key], value)
        else:
            merged[key] = value
    return merged


def zip_files(output_zip_path: str, *paths: str) -> None:
    """Directories keep their own name as the top-level entry (``model/...``)."""
    with zipfile.ZipFile(output_zip_path, "w", zipfile.ZIP_DEFLATED) as zf:
        for path in paths:
            if os.path.isfile(path):
                zf.write(path, arcname=os.path.basename(path))
            elif os.path.isdir(path):
                parent = os.path.dirname(os.path.normpath(path))
                for root, _, files in os.walk(path):
                    for name in sorted(files):
                        full = os.path.join(root, name)
                        zf.write(full, arcname=os.path.relpath(full, parent))
            else:
                logger.warning(f"Skipped: Path '{path}' does not exist.")


def unzip_files(zip_path: str, output_folder: str) -> None:
    if not os.path.isfile(zip_path):
        logger.error(f"ZIP file '{zip_path}' does not exist.")
        return
    os.makedirs(output_folder, exist_ok=True)
    with zipfile.ZipFile(zip_path, "r") as zf:
        zf.extractall(output_folder)


def files_exist(*paths: str) -> bool:
    return all(os.path.isfile(p) for p in paths)
