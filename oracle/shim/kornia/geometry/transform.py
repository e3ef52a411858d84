def hflip(x):
    return x.flip(-1)


def vflip(x):
    return x.flip(-2)
