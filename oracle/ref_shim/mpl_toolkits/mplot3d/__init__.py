"""TEST INFRASTRUCTURE: import stub."""


class Axes3D(object):
    pass
