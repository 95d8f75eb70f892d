"""Frame egress compatible with the reference's sinks: host I/O, not accelerated.

``message_socket_sink(ip, port, layer)`` sends every frame published by a decoder as one UDP
datagram, stripped according to ``layer`` exactly like lib/message_socket_sink_impl.cc:93-122:
LORATAP (0) = whole blob, LORAPHY (1) = without the 15-byte loratap header, LORAMAC (2) = payload
only, without the PHY header and without the two MAC CRC bytes when the header says they exist.
``LoRaUDPServer`` is the receiving end the reference's test-suite uses (python/lorasocket.py:4-34):
with both, python/qa_testsuite.py's scoring method (hex payload equality over UDP, :104-125,244-245)
runs against the GPU decoder."""
from __future__ import annotations

import binascii
import socket

from .decoder import LORAPHY_LEN, LORATAP_LEN

LORATAP, LORAPHY, LORAMAC = 0, 1, 2            # include/lora/message_socket_sink.h:695
MAC_CRC_SIZE = 2                               # include/lora/utilities.h:29


def strip_layers(blob: bytes, layer: int) -> bytes:
    """msg_send_udp, lib/message_socket_sink_impl.cc:93-116."""
    if layer == LORAPHY:
        return blob[LORATAP_LEN:]
    if layer == LORAMAC:
        has_mac_crc = (blob[LORATAP_LEN + 1] >> 4) & 1          # loraphy_header_t.has_mac_crc
        end = len(blob) - MAC_CRC_SIZE * has_mac_crc
        return blob[LORATAP_LEN + LORAPHY_LEN:end]
    return blob


class message_socket_sink:
    def __init__(self, ip="127.0.0.1", port=40868, layer=LORATAP):
        self.addr, self.layer = (ip, int(port)), int(layer)
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)

    def handle(self, stream: int, blob: bytes):        # signature of decoder.message_port_subscribe handlers
        self.sock.sendto(strip_layers(blob, self.layer), self.addr)

    def connect(self, block):
        """msg_connect((block, 'frames'), (self, 'in'))"""
        block.message_port_subscribe(self.handle)
        return self

    def close(self):
        self.sock.close()


class LoRaUDPServer:
    def __init__(self, ip="127.0.0.1", port=40868, timeout=10):
        self.s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.s.bind((ip, port))
        self.s.settimeout(timeout)

    def get_payloads(self, number_of_payloads):
        out = []
        for _ in range(number_of_payloads):
            try:
                data = self.s.recvfrom(65535)[0]
                if data:
                    out.append(binascii.hexlify(data))
            except Exception as exc:        # the reference prints and carries on
                print(exc)
        return out

    def close(self):
        self.s.close()
