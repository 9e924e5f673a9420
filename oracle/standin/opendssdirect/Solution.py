def Solve():
    return None
