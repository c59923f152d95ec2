#!/usr/bin/env python
"""Chunked reply transport for the reference's server / client.

The unchanged reply path caps a usable reply at a few hundred KB:
  * src/server.c:521-524 copies the whole reply into a stack VLA (`char send_buffer[length + 1]`,
    never read again) -- a multi-MB `print` overflows the stack;
  * src/client.c:126-133 receives the reply into another stack VLA with ONE recv(), so anything
    beyond what the socket buffer held at that moment is dropped and then misread as the next
    header.
The wire format stays what it is -- a `message` header carrying the byte length, then the
payload -- but the payload now travels in pieces: the server loops over send() until every byte
is out (no copy, no VLA), the client mallocs `length + 1` bytes and loops over recv() until it
has them all.  Replies up to INT_MAX bytes (the header's `int length`).

This script writes PATCHED COPIES of the two files for the drop-in build (oracle/Makefile puts
them under oracle/_ref/dropin/patched/, git-ignored); the reference tree is never modified and
no reference source is committed -- the edits are located by the statements they replace.
INTEGRATION.md shows the same change as a diff for a maintainer.

usage: apply_reply_patch.py <reference src dir> <output dir>
"""
import os
import re
import sys


def patch_server(text: str) -> str:
    # (1) drop the stack copy of the reply
    vla = re.compile(r"[ \t]*char send_buffer\[send_message\.length \+ 1\];\s*\n"
                     r"[ \t]*strcpy\(send_buffer, result\);\s*\n"
                     r"[ \t]*send_message\.payload = send_buffer;\s*\n")
    text, n = vla.subn("            send_message.payload = result;           /* no stack copy of the reply */\n", text)
    assert n == 1, "server.c: reply VLA not found"
    # (2) send the payload until every byte is out
    one_send = re.compile(r"[ \t]*if \(send\(client_socket, result, send_message\.length, 0\) == -1\) \{\s*\n"
                          r"[ \t]*log_err\(\"Failed to send message\.\"\);\s*\n"
                          r"[ \t]*exit\(1\);\s*\n"
                          r"[ \t]*\}\s*\n")
    loop = ("            for (size_t adb_sent = 0; adb_sent < (size_t)send_message.length;) {\n"
            "                size_t adb_piece = (size_t)send_message.length - adb_sent;\n"
            "                if (adb_piece > ((size_t)1 << 20)) adb_piece = (size_t)1 << 20;\n"
            "                ssize_t adb_n = send(client_socket, result + adb_sent, adb_piece, 0);\n"
            "                if (adb_n <= 0) {\n"
            "                    log_err(\"Failed to send message.\");\n"
            "                    exit(1);\n"
            "                }\n"
            "                adb_sent += (size_t)adb_n;\n"
            "            }\n")
    text, n = one_send.subn(lambda m: loop, text)
    assert n == 1, "server.c: payload send not found"
    return text


def patch_client(text: str) -> str:
    block = re.compile(r"[ \t]*char payload\[num_bytes \+ 1\];\s*\n"
                       r"(?:[ \t]*\n|[ \t]*//[^\n]*\n)*"
                       r"[ \t]*if \(\(len = recv\(client_socket, payload, num_bytes, 0\)\) > 0\) \{\s*\n"
                       r"[ \t]*payload\[num_bytes\] = '\\0';\s*\n"
                       r"[ \t]*printf\(\"%s\\n\", payload\);\s*\n"
                       r"[ \t]*\}\s*\n")
    loop = ("                    char *payload = malloc((size_t)num_bytes + 1);     /* heap, not a stack VLA */\n"
            "                    int adb_got = 0;\n"
            "                    while (payload && adb_got < num_bytes &&\n"
            "                           (len = recv(client_socket, payload + adb_got, num_bytes - adb_got, 0)) > 0)\n"
            "                        adb_got += len;                               /* until the whole reply is in */\n"
            "                    if (payload && adb_got == num_bytes) {\n"
            "                        payload[num_bytes] = '\\0';\n"
            "                        fwrite(payload, 1, (size_t)num_bytes, stdout);\n"
            "                        fputc('\\n', stdout);\n"
            "                    }\n"
            "                    free(payload);\n")
    text, n = block.subn(lambda m: loop, text)
    assert n == 1, "client.c: reply receive not found"
    return text


def main():
    src, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for name, fn in (("server.c", patch_server), ("client.c", patch_client)):
        with open(os.path.join(src, name)) as f:
            text = f.read()
        with open(os.path.join(out, name), "w") as f:
            f.write(fn(text))


if __name__ == "__main__":
    main()
