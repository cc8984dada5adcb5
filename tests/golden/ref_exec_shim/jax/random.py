def PRNGKey(seed):
    return int(seed)


def split(key, num=2):
    return [key + i + 1 for i in range(num)]
