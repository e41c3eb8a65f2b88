class BotorchError(Exception):
    pass


class BotorchTensorDimensionError(BotorchError):
    pass


class UnsupportedError(BotorchError):
    pass
