"""Action <-> move <-> string codec (chess_v2.py:492-567, lib.rs:1278-1373): table lookups of the product package against
the oracle's codec and the literal examples of the reference."""
from gym_chess_b200 import codec
from oracle import oracle as orc


def test_tables_roundtrip_and_match_the_oracle():
    assert codec.NUM_ACTIONS == 4101 and len(codec.ACTION_TO_STR) == 4101 and len(codec.ACTION_TO_MOVE) == 4101
    for a in range(4100):
        s = codec.ACTION_TO_STR[a]
        assert s == orc.action_to_str(a) and orc.str_to_action(s) == a and codec.STR_TO_ACTION[s] == a
        m = codec.action_to_move(a)
        assert codec.move_to_action(m) == a and codec.move_to_str_code(m) == s and codec.str_code_to_move(s) == m
    assert codec.ACTION_TO_STR[4100] == "RESIGN" and codec.move_to_action("RESIGN") == 4100


def test_reference_examples():
    # chess_v2.py:492-506: action = (r0*8+c0)*64 + (r1*8+c1); lib.rs:1278-1290: file letter + (8 - row)
    assert codec.move_to_action(((6, 4), (4, 4))) == 3364 and codec.ACTION_TO_STR[3364] == "e2e4"
    assert codec.action_to_move(3112) == ((6, 0), (5, 0)) and codec.ACTION_TO_STR[3112] == "a2a3"   # first start-position move
    assert [codec.move_to_action(m) for m in codec.CASTLE_MOVES] == [4096, 4097, 4098, 4099]
    assert codec.ACTION_FROM[3364] == 52 and codec.ACTION_TO[3364] == 36 and codec.ACTION_FROM[4097] == -1
