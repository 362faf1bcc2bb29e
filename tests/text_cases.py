"""Strings for the tokeniser parity tests, written with ASCII escapes only: every whitespace character of str.isspace()
appears, plus multi-byte characters that are NOT whitespace."""

DOCS = [
    "Le tribunal  de première instance\test compétent",
    "article 5 : « nul n'est censé ignorer la loi » !",       # NBSP / narrow NBSP
    "   leading and trailing   ",
    "",
    "tab\tnew\nline\rcarriage\x0bvt\x0cff\x1cfs\x1dgs\x1ers\x1fus",
    "narrow nbsp thin space ideographic　space line sep para sep ogham mark math space nelx",
    "en quad em quad        hair",
    "é è ê ë ç à ù œ … — « » ’",   # multi-byte non-space characters
    "zero​width is NOT whitespace for str.split",
    "emoji \U0001f600 and 汉字 mixed 汉字",
    "a a a b a",
]

QUERIES = ["tribunal compétent inconnu", "汉字 \U0001f600 zero​width", "", "loi » a"]
