"""tpgx — B200-native evaluator for Tangled Program Graph policies (Python-side binding of the C ABI)."""
from . import _cabi as abi
from .api import Evaluator, TpgxError, load_library, LIB_PATH, EXPORTS, mnist_episode_indices, load_idx, flatten_programs, slice_graph, k1v2_encode_program, k1v2_handler_names, kernel_launches, comm_init_all, evaluate_classification_allgather_all
from .graph import (TPGraph, INSTRUCTION_SETS, APP_SOURCES, APP_CONSTANTS, APP_ACTIONS, APP_ENV, import_dot, export_dot,
                    synthetic_graph, synthetic_images, random_programs, address_space, LINE_DTYPE)
from .train import Trainer, load_params_json, python_backend, episodes_backend, classification_backend
from .sharding import shard_range, slot_shard, gather_records, records_from
