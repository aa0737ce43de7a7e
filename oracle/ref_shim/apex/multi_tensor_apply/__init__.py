def multi_tensor_applier(op, noop_flag, tensor_lists, *args):
    return op(2048 * 32, noop_flag, tensor_lists, *args)
