-- Rotor-angle sweep of the TorqueBenchmark problem through the reference's Lua console: every position is saved,
-- meshed, analysed (mi_analyze) and post-processed (mo_gapintegral) by femmcli; prints "<degrees> <torque>".
-- tests/test_gpu_femmcli.py runs it with oracle/_ref/femmcli (CPU reference) and oracle/_ref/femmcli_b200 (the same
-- console with mi_analyze on the GPU) and checks the torques against the analytic sin(angle) N*m.
open("tb.fem")
for deg = 0, 90, 10 do
	mi_modifyboundprop("AGE", 10, deg)
	mi_modifyboundprop("AGE", 11, 0)
	mi_saveas("tb_" .. deg .. ".fem")
	mi_createmesh()
	mi_analyze()
	mi_loadsolution()
	write(format("TORQUE %d %.12g\n", deg, mo_gapintegral("AGE", 0)))
end
write("DONE\n")
