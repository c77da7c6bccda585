"""speechbrain.utils.checkpoints stand-in (only the key-rename helper used at Transformer.py:865-869)."""


def map_old_state_dict_weights(state_dict, mapping):
    for old, new in mapping.items():
        for key in list(state_dict.keys()):
            if old in key:
                state_dict[key.replace(old, new)] = state_dict.pop(key)
    return state_dict
