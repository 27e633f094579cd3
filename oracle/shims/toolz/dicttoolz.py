def merge_with(func, *dicts):
    """merge_with(f, d1, d2, ...) -> {k: f([d1[k], d2[k], ...])} for keys present."""
    if len(dicts) == 1 and not isinstance(dicts[0], dict):
        dicts = tuple(dicts[0])
    acc = {}
    for d in dicts:
        for k, v in d.items():
            acc.setdefault(k, []).append(v)
    return {k: func(v) for k, v in acc.items()}
