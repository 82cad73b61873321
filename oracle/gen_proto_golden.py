"""Generate tests/golden/generation_proto_schema.json from the REFERENCE's wire schema
(/root/reference/src/vllm_tgis_adapter/grpc/pb/generation.proto), in the build container:  python oracle/gen_proto_golden.py

The repo has no protoc, so vllm_tgis_adapter_b200/grpc/pb/generation_pb2.py builds the descriptor by hand; this golden
(every message / field name, number, type, label, proto3-optional flag, oneof membership, enum value, and RPC signature
of the reference .proto) is what tests/test_grpc_server_cpu.py::test_proto_descriptor_matches_reference_schema pins it
to.  The parser below understands exactly the proto3 subset that file uses (nested messages/enums, `optional`, `repeated`,
`oneof`, `rpc ... returns (stream ...)`)."""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/src/vllm_tgis_adapter/grpc/pb/generation.proto")
OUT = ROOT / "tests" / "golden" / "generation_proto_schema.json"


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def tokenize(text: str) -> list[str]:
    return re.findall(r"[A-Za-z_][A-Za-z0-9_.]*|\d+|\"[^\"]*\"|[{}()=;<>,\[\]]", text)


class Parser:
    def __init__(self, toks: list[str]):
        self.t, self.i = toks, 0
        self.schema = {"package": None, "syntax": None, "messages": {}, "enums": {}, "services": {}}

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        v = self.t[self.i]
        self.i += 1
        return v

    def expect(self, v):
        got = self.next()
        assert got == v, (got, v, self.t[max(0, self.i - 5): self.i + 5])

    def parse(self):
        while self.peek() is not None:
            k = self.next()
            if k == "syntax":
                self.expect("=")
                self.schema["syntax"] = self.next().strip('"')
                self.expect(";")
            elif k == "package":
                self.schema["package"] = self.next()
                self.expect(";")
            elif k == "option" or k == "import":
                while self.next() != ";":
                    pass
            elif k == "message":
                self.message("")
            elif k == "enum":
                self.enum("")
            elif k == "service":
                self.service()
            else:
                raise AssertionError(f"unexpected top-level token {k}")
        return self.schema

    def enum(self, prefix: str):
        name = prefix + self.next()
        self.expect("{")
        vals = {}
        while self.peek() != "}":
            n = self.next()
            self.expect("=")
            vals[n] = int(self.next())
            self.expect(";")
        self.expect("}")
        self.schema["enums"][name] = vals

    def message(self, prefix: str):
        name = prefix + self.next()
        self.expect("{")
        fields: dict = {}
        self.schema["messages"][name] = fields
        self.body(name, fields, None)

    def body(self, name: str, fields: dict, oneof: str | None):
        while self.peek() != "}":
            k = self.next()
            if k == "message":
                self.message(name + ".")
            elif k == "enum":
                self.enum(name + ".")
            elif k == "oneof":
                on = self.next()
                self.expect("{")
                self.body(name, fields, on)
            elif k == "reserved":
                while self.next() != ";":
                    pass
            else:
                label = "singular"
                if k in ("optional", "repeated"):
                    label = k
                    k = self.next()
                ftype = k
                fname = self.next()
                self.expect("=")
                num = int(self.next())
                if self.peek() == "[":
                    while self.next() != "]":
                        pass
                self.expect(";")
                fields[fname] = {"number": num, "type": ftype, "label": label, "oneof": oneof}
        self.expect("}")

    def service(self):
        name = self.next()
        self.expect("{")
        rpcs = {}
        while self.peek() != "}":
            self.expect("rpc")
            rn = self.next()
            self.expect("(")
            cs = self.peek() == "stream"
            if cs:
                self.next()
            inp = self.next()
            self.expect(")")
            self.expect("returns")
            self.expect("(")
            ss = self.peek() == "stream"
            if ss:
                self.next()
            outp = self.next()
            self.expect(")")
            if self.peek() == "{":
                self.next()
                self.expect("}")
            else:
                self.expect(";")
            rpcs[rn] = {"input": inp, "output": outp, "client_streaming": cs, "server_streaming": ss}
        self.expect("}")
        self.schema["services"][name] = rpcs


def main() -> None:
    schema = Parser(tokenize(strip_comments(SRC.read_text()))).parse()
    schema["meta"] = {"generated_by": "oracle/gen_proto_golden.py", "source": str(SRC),
                      "n_fields": sum(len(f) for f in schema["messages"].values()),
                      "n_enum_values": sum(len(e) for e in schema["enums"].values())}
    OUT.write_text(json.dumps(schema, indent=1, sort_keys=True))
    print("wrote", OUT, schema["meta"])


if __name__ == "__main__":
    sys.exit(main())
