"""mpyc.sectypes look-alike: only the class names the reference uses in isinstance() checks."""


class SecureObject:
    pass


class SecureNumber(SecureObject):
    pass


class SecureFiniteField(SecureNumber):
    pass


class SecureInteger(SecureNumber):
    pass
