/* TEST INFRASTRUCTURE: a 20-line main around the reference's own UDP sender (src/utils/net.c new_socket,
 * socket_sender_bind, socket_send -- compiled from the reference tree in place by oracle/Makefile `ref`, never copied):
 * sends the bytes of a file as one datagram.  tests/test_ingest.py uses it to show that the ingest receiver
 * interoperates with the reference's socket layer.   usage: net_send <addr> <port> <file> */
#include <stdio.h>
#include <stdlib.h>
#include "utils/net.h"

int main(int argc, char **argv) {
  if (argc != 4) return 2;
  FILE *f = fopen(argv[3], "rb");
  if (!f) return 3;
  static char buf[65536];
  const int n = (int)fread(buf, 1, sizeof buf, f);
  fclose(f);
  struct Socket *s = new_socket(atoi(argv[2]), argv[1]);
  if (!socket_sender_bind(s)) return 4;
  const int sent = socket_send(s, buf, n);
  delete_socket(s);
  return sent == n ? 0 : 5;
}
