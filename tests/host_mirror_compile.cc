// compile-only check of the C++ host mirror (tests/test_capi_symbols.py builds it with g++)
#include "../alfalfa_b200/host/alfalfa_gpu.hh"
int use(const uint8_t* data, uint64_t n) {
  alfalfa_gpu::Context ctx(0, 320, 240);
  alfalfa_gpu::Decoder dec(ctx, 320, 240);
  alfalfa_gpu::Decoder copy = dec;
  auto out = dec.get_frame_output(alfalfa_gpu::Chunk(data, n));
  auto parsed = copy.parse_frame(alfalfa_gpu::Chunk(data, n));
  auto out2 = copy.decode_frame(parsed);
  alfalfa_gpu::Encoder enc(ctx, 320, 240);
  alfalfa_gpu::SourceFrame sf = {data, data, data, 320, 160};
  auto bytes = enc.encode_with_target_size(sf, 20000);
  bytes = enc.encode_with_quantizer(sf, 40);
  auto rec = enc.reconstruction();
  (void)rec;
  return (int)bytes.size() + (dec == copy) + out.first + out2.first + (int)out.second.dump(320, 240).size() +
         (dec.get_state() == copy.get_state()) + (dec.get_references().last == copy.get_references().last);
}
