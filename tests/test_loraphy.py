"""Host-side frame view (N4): PHY header fields and the 5-bit header checksum the reference never verifies."""
import pytest

from gr_lora_b200 import tx
from gr_lora_b200.loraphy import parse_frame


def frame(header: bytes, payload: bytes, snr: int = 0) -> bytes:
    tap = bytearray(15)
    tap[13] = snr
    return bytes(tap) + header + payload


def test_readme_golden_header_checks_out():
    # README.md:67-71: " 04 90 40 de ad be ef 70 0d"
    f = parse_frame(frame(bytes.fromhex("049040"), bytes.fromhex("deadbeef700d"), snr=7))
    assert (f.length, f.cr, f.has_mac_crc, f.checksum, f.header_ok) == (4, 4, True, 0b00100, True)
    assert f.payload.hex() == "deadbeef700d" and f.snr == 7


@pytest.mark.parametrize("length,cr,crc", [(0, 1, 0), (1, 2, 1), (16, 3, 0), (200, 4, 1), (255, 1, 1)])
def test_every_single_bit_error_in_the_header_is_caught(length, cr, crc):
    h = tx.header_bytes(length, cr, crc)
    assert parse_frame(frame(h, b"")).header_ok
    for bit in range(20):                       # the 20 used bits: 8 length + 3 cr + 1 crc + 5 checksum... all of b0, b1, b2[7:4]
        byte, pos = divmod(bit, 8)
        if byte == 1 and pos in (1, 2, 3):      # crc_msn bits 1..3 are not part of the checksum field
            continue
        bad = bytearray(h)
        bad[byte] ^= 1 << pos if byte < 2 else 1 << (4 + pos)
        assert not parse_frame(frame(bytes(bad), b"")).header_ok, (bit, bad.hex())


def test_short_blob_is_rejected():
    with pytest.raises(ValueError):
        parse_frame(b"\x00" * 17)
