"""CPU: frame egress layers (lib/message_socket_sink_impl.cc:93-122) and the UDP loop the reference's
qa_testsuite scores with (python/lorasocket.py, python/qa_testsuite.py:104-125)."""
import socket

from gr_lora_b200.message_socket_sink import LORAMAC, LORAPHY, LORATAP, LoRaUDPServer, message_socket_sink, strip_layers

FRAME = bytes(13) + b"\x2c\x00" + bytes.fromhex("049040deadbeef700d")     # loratap(15, snr=0x2c) | phy | payload+crc


def test_strip_layers():
    assert strip_layers(FRAME, LORATAP) == FRAME
    assert strip_layers(FRAME, LORAPHY).hex() == "049040deadbeef700d"
    assert strip_layers(FRAME, LORAMAC).hex() == "deadbeef"                # has_mac_crc = 1: CRC bytes dropped
    no_crc = bytes(15) + bytes.fromhex("0480408899aabb")
    assert strip_layers(no_crc, LORAMAC).hex() == "8899aabb"


def test_udp_round_trip_like_qa_testsuite():
    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    server = LoRaUDPServer(port=port, timeout=5)
    sink = message_socket_sink("127.0.0.1", port, LORAMAC)

    class FakeBlock:                                    # anything with the decoder's message-port interface
        def message_port_subscribe(self, h):
            self.h = h
    blk = FakeBlock()
    sink.connect(blk)
    for _ in range(3):
        blk.h(0, FRAME)
    assert server.get_payloads(3) == [b"deadbeef"] * 3   # what qa_testsuite compares with test:expected
    sink.close()
    server.close()
