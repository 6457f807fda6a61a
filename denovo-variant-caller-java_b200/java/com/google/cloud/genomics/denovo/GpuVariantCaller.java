/*
 * GpuVariantCaller.java — drop-in for VariantCaller (VariantCaller.java:55-210).  SOURCE ONLY.
 * The (Variant, VariantCall) stream of one position shard is flattened into dnv_vcall records — exactly
 * the fields VariantsBuffer reads (VariantsBuffer.java:172-306) — and ONE downcall replaces the
 * checkAndAdd / canProcess / retrieveNextCall loop; denovo PositionCalls come back in stream order.
 */
package com.google.cloud.genomics.denovo;

import com.google.api.services.genomics.model.Variant;
import com.google.api.services.genomics.model.VariantCall;
import com.google.cloud.genomics.denovo.DenovoUtil.Genotype;
import com.google.cloud.genomics.denovo.DenovoUtil.TrioMember;
import com.google.cloud.genomics.denovo.gpu.DenovoB200;

import java.io.PrintWriter;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.util.List;

public class GpuVariantCaller {

  private final DenovoShared shared;
  private final DenovoB200 engine;

  public GpuVariantCaller(DenovoShared shared, DenovoB200 engine) {
    this.shared = shared;
    this.engine = engine;
  }

  /** Flatten one (variant, call) pair (see dnv_vcall in include/denovo_b200.h). */
  static void pack(MemorySegment rec, Variant variant, VariantCall call, TrioMember person, int variantId) {
    List<Integer> gt = call.getGenotype();
    String ref = variant.getReferenceBases();
    rec.set(ValueLayout.JAVA_LONG, 0, variant.getStart());
    rec.set(ValueLayout.JAVA_LONG, 8, variant.getEnd());
    rec.set(ValueLayout.JAVA_INT, 16, variantId);
    rec.set(ValueLayout.JAVA_BYTE, 20, (byte) (gt.size() > 0 ? gt.get(0) : 0));
    rec.set(ValueLayout.JAVA_BYTE, 21, (byte) (gt.size() > 1 ? gt.get(1) : 0));
    rec.set(ValueLayout.JAVA_BYTE, 22, (byte) Math.min(gt.size(), 255));
    rec.set(ValueLayout.JAVA_BYTE, 23, (byte) person.ordinal()); // CHILD 0, MOM 1, DAD 2
    List<String> filter = call.getInfo() == null ? null : call.getInfo().get("FILTER");
    boolean pass = filter != null && filter.size() == 1 && filter.get(0).equals("PASS"); // VariantsBuffer:245-249
    rec.set(ValueLayout.JAVA_BYTE, 24, (byte) (pass ? 1 : 0));
    rec.set(ValueLayout.JAVA_BYTE, 25, (byte) Math.min(ref.length(), 255));
    rec.set(ValueLayout.JAVA_BYTE, 26, (byte) (ref.isEmpty() ? 0 : ref.charAt(0)));
    for (int k = 0; k < 2; k++) {
      int g = gt.size() > k ? gt.get(k) : 0;
      String alt = g > 0 ? variant.getAlternateBases().get(g - 1) : "";
      rec.set(ValueLayout.JAVA_BYTE, 27 + 2L * k, (byte) Math.min(alt.length(), 255));
      rec.set(ValueLayout.JAVA_BYTE, 28 + 2L * k, (byte) (alt.isEmpty() ? 0 : alt.charAt(0)));
    }
  }

  /** callSimpleDenovo over an already-fetched, start-sorted shard. */
  void callShard(PrintWriter callWriter, String contig, List<Variant> variants, DenovoCaller writer) {
    long n = 0;
    for (Variant v : variants) if (v.getCalls() != null) n += v.getCalls().size(); // VariantCaller:147-153
    if (n == 0) return;
    try (Arena arena = Arena.ofConfined()) {
      MemorySegment recs = arena.allocate(n * DenovoB200.VCALL.byteSize(), 8);
      long i = 0;
      int vid = 0;
      for (Variant v : variants) {
        if (v.getCalls() == null) continue;
        for (VariantCall call : v.getCalls()) {
          pack(recs.asSlice(i * 32, 32), v, call, shared.getCallsetIdToPersonMap().get(call.getCallSetId()), vid);
          i++;
        }
        vid++;
      }
      MemorySegment out = arena.allocate(n * DenovoB200.PCALL.byteSize(), 8);
      long nOut = engine.mendelFilter(recs, n, out, n);
      Genotype[] g = Genotype.values();
      StringBuilder sb = new StringBuilder();
      for (long k = 0; k < nOut; k++) {
        MemorySegment c = out.asSlice(k * 24, 24);
        if (c.get(ValueLayout.JAVA_BYTE, 19) == 0) continue; // is_denovo
        long pos = c.get(ValueLayout.JAVA_LONG, 0);
        // PositionCall.toString, VariantsBuffer.java:358-360
        sb.append(String.format("%s,%d,[%d,{CHILD=%s, MOM=%s, DAD=%s}]%n", contig, pos, pos,
            g[c.get(ValueLayout.JAVA_BYTE, 16)], g[c.get(ValueLayout.JAVA_BYTE, 17)], g[c.get(ValueLayout.JAVA_BYTE, 18)]));
      }
      writer.writeCalls(callWriter, sb.toString());
    }
  }
}
