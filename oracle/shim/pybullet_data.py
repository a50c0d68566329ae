def getDataPath():
    raise NotImplementedError('pybullet_data is not available (oracle shim)')
