"""Bulk-ingest feeders: the CSV / JSONL loaders of the reference
(wdbx/utils/data_utils.py:16-231, same signatures and return shape) plus ``ingest_*`` helpers that
push a whole file through ``VectorStore.batch_store`` -- one stacked H2D copy + one ingest kernel (K4)
per shard instead of one Python call per row.
"""
from __future__ import annotations

import csv
import json
import logging
from typing import Any, Dict, List, Optional, Tuple, Union

logger = logging.getLogger(__name__)

Vectors = Dict[str, List[float]]
Metadata = Dict[str, Dict[str, Any]]


def parse_vector(vector_data: Union[str, List, Dict]) -> List[float]:
    """list | '[1, 2]' | '1,2' | '1 2' | '[1. 2.]' (numpy str) | {'vector'|'embedding'|'values'|'data': ...}
    -> list of floats.  Same acceptance set and the same exception types as the reference (data_utils.py:174-231;
    tests/golden/loader_golden.json): a JSON array whose elements are not numbers raises what ``float`` raises
    (only a JSON syntax error falls through to the other notations), and ``'array([1., 2.])'`` -- numpy's repr, with
    commas -- is NOT accepted, exactly as there."""
    if isinstance(vector_data, list):
        return [float(x) for x in vector_data]
    if isinstance(vector_data, dict):
        for field in ("vector", "embedding", "values", "data"):
            if field in vector_data:
                return parse_vector(vector_data[field])
        raise ValueError(f"Could not find vector data in dictionary: {vector_data}")
    if not isinstance(vector_data, str):
        raise ValueError(f"Unsupported vector data type: {type(vector_data)}")
    text = vector_data.strip()
    if text.startswith("[") and text.endswith("]"):
        try:
            parsed = json.loads(text)
        except json.JSONDecodeError:
            parsed = None
        else:
            return [float(x) for x in parsed]
    stripped = text.replace("array(", "").replace(")", "").replace("[", "").replace("]", "")
    for pieces in (lambda: text.split(","), lambda: text.split(), lambda: stripped.split()):
        try:
            return [float(x.strip()) for x in pieces()]
        except ValueError:
            pass
    raise ValueError(f"Could not parse vector from string: {vector_data}")


def load_vectors_from_csv(file_path: str, vector_column: Union[str, int], id_column: Optional[Union[str, int]] = None,
                          delimiter: str = ",", skip_header: bool = True,
                          metadata_columns: Optional[List[Union[str, int]]] = None) -> Tuple[Vectors, Metadata]:
    """Reference: data_utils.py:16-108.  Column names -> DictReader (ids default to ``row_{i}``); column
    indices -> plain reader (metadata keys ``col_{j} ``, as the reference writes them)."""
    vectors: Vectors = {}
    metadata: Metadata = {}
    by_name = isinstance(vector_column, str) or bool(metadata_columns and any(isinstance(c, str) for c in metadata_columns))
    try:
        with open(file_path, "r", encoding="utf-8") as f:   # universal newlines, as the reference opens it
            if by_name:
                rows = csv.DictReader(f, delimiter=delimiter)
            else:
                rows = csv.reader(f, delimiter=delimiter)
                if skip_header:
                    next(rows, None)
            for i, row in enumerate(rows):
                try:
                    if by_name:
                        vid = row[id_column] if id_column else f"row_{i}"
                        vectors[vid] = parse_vector(row[vector_column])
                        metadata[vid] = {c: row[c] for c in (metadata_columns or []) if c in row}
                    else:
                        vid = row[id_column] if id_column is not None else f"row_{i}"
                        vectors[vid] = parse_vector(row[vector_column])
                        metadata[vid] = {f"col_{j} ": row[j] for j in (metadata_columns or []) if j < len(row)}
                except Exception as e:  # a bad row is skipped, like the reference does
                    logger.warning(f"Error processing row {i}: {e}")
    except Exception as e:
        logger.error(f"Error loading vectors from CSV {file_path}: {e}")
        raise ValueError(f"Error loading vectors from CSV: {e}")
    return vectors, metadata


def load_vectors_from_jsonl(file_path: str, vector_field: str, id_field: Optional[str] = None,
                            metadata_fields: Optional[List[str]] = None) -> Tuple[Vectors, Metadata]:
    """Reference: data_utils.py:111-171 (ids default to ``line_{i}``; all other fields become metadata
    unless ``metadata_fields`` is given)."""
    vectors: Vectors = {}
    metadata: Metadata = {}
    try:
        with open(file_path, "r", encoding="utf-8") as f:
            for i, line in enumerate(f):
                try:
                    obj = json.loads(line.strip())
                    vid = obj.get(id_field) if id_field else f"line_{i}"
                    if vector_field not in obj:
                        logger.warning(f"Vector field '{vector_field}' not found in line {i}")
                        continue
                    vectors[vid] = parse_vector(obj[vector_field])
                    if metadata_fields:
                        metadata[vid] = {k: obj[k] for k in metadata_fields if k in obj}
                    else:
                        metadata[vid] = {k: v for k, v in obj.items() if k != vector_field}
                except Exception as e:
                    logger.warning(f"Error processing line {i}: {e}")
    except Exception as e:
        logger.error(f"Error loading vectors from JSONL {file_path}: {e}")
        raise ValueError(f"Error loading vectors from JSONL: {e}")
    return vectors, metadata


def _ingest(store, vectors: Vectors, metadata: Metadata) -> int:
    dim = store.vector_dim
    good = {vid: v for vid, v in vectors.items() if len(v) == dim}
    if len(good) != len(vectors):
        logger.warning(f"skipped {len(vectors) - len(good)} vectors whose dimension is not {dim}")
    return store.batch_store(good, {vid: metadata.get(vid, {}) for vid in good})


def ingest_csv(store, file_path: str, vector_column, id_column=None, delimiter: str = ",", skip_header: bool = True,
               metadata_columns=None) -> int:
    """Load a CSV file and store it through ``VectorStore.batch_store`` (one device append per shard)."""
    return _ingest(store, *load_vectors_from_csv(file_path, vector_column, id_column, delimiter, skip_header, metadata_columns))


def ingest_jsonl(store, file_path: str, vector_field: str, id_field=None, metadata_fields=None) -> int:
    """Load a JSONL file and store it through ``VectorStore.batch_store``."""
    return _ingest(store, *load_vectors_from_jsonl(file_path, vector_field, id_field, metadata_fields))
